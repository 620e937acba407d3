"""mnk_b200 -- B200 (sm_100a) implementation of the rl-selfplay-mnk hot path.

Host side: PyTorch tensors for memory / streams / torch.distributed; compute: hand-written CUDA
kernels in libmnk_b200.so behind the C ABI of include/mnk_b200.h.  The sibling packages ``env/``
and ``selfplay/`` re-export these classes under the reference's module paths so that putting
``rl-selfplay-mnk_b200/`` ahead of the reference's ``src/`` on sys.path swaps the path in.
"""
from ._lib import build, lib, LIB_PATH  # noqa: F401
from .env import TorchVectorMnkEnv  # noqa: F401
from .policy import NNPolicy, Policy, RandomPolicy  # noqa: F401
from .sampling import MaskedCategorical, masked_sample  # noqa: F401
from .wrapper import TorchSelfPlayWrapper  # noqa: F401
from .rollout import RolloutBuffer, RolloutCollector, RolloutStats  # noqa: F401
from .nets import CnnActorCritic, ResNetActorCritic, TransformerActorCritic, build_architecture  # noqa: F401
from .resnet import NativeNNPolicy, NativeResNet  # noqa: F401
from .convnet import NativeConvNet, native_network  # noqa: F401
from .transformer import NativeTransformer  # noqa: F401
from .ppo import PPOAgent, TrainingMetrics  # noqa: F401
from . import dist, model_io  # noqa: F401

__all__ = ["build", "lib", "LIB_PATH", "TorchVectorMnkEnv", "TorchSelfPlayWrapper", "Policy", "RandomPolicy", "NNPolicy",
           "MaskedCategorical", "masked_sample", "RolloutBuffer", "RolloutCollector", "RolloutStats", "dist", "model_io",
           "ResNetActorCritic", "CnnActorCritic", "build_architecture", "NativeResNet", "NativeConvNet", "NativeTransformer", "TransformerActorCritic", "native_network",
           "NativeNNPolicy", "PPOAgent", "TrainingMetrics"]
