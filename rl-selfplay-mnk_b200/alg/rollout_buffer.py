"""Drop-in module path for the reference's ``alg.rollout_buffer`` (src/alg/ppo.py:8): the packed
bitboard buffer of mnk_b200.rollout with the reference's interface."""
from mnk_b200.rollout import RolloutBuffer  # noqa: F401
