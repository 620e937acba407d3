"""Thin alias of oracle.ref_tree (kept for the tests that import it): the unmodified reference, mounted at
/root/reference in the build container or staged under oracle/_ref/ by ``__graft_entry__.build()``."""
from oracle.ref_tree import available, load, load_dropin, root  # noqa: F401
