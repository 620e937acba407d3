"""Drop-in package ``env``: modules defined here replace the reference's src/env/ modules of the same
name; the others are taken from the reference's src/env/ when it is on sys.path (mnk_b200._overlay)."""
from mnk_b200 import _overlay

_overlay.extend(__name__, __path__)
