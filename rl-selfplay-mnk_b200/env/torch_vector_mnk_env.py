"""Drop-in module path for the reference's ``env.torch_vector_mnk_env`` (src/train.py:10,
src/selfplay/validation.py:2): the class is the sm_100a implementation in mnk_b200.env."""
from mnk_b200.env import TorchVectorMnkEnv  # noqa: F401
from .constants import PLAYER_BLACK, PLAYER_WHITE  # noqa: F401
