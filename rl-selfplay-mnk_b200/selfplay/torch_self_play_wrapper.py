"""Drop-in module path for the reference's ``selfplay.torch_self_play_wrapper`` (src/train.py:11,
src/selfplay/validation.py:3): the class is the fused-kernel implementation in mnk_b200.wrapper."""
from mnk_b200.wrapper import TorchSelfPlayWrapper  # noqa: F401
