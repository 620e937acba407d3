"""Drop-in module path for the reference's ``selfplay.policy`` (src/train.py:12)."""
from mnk_b200.policy import NNPolicy, Policy, RandomPolicy  # noqa: F401
