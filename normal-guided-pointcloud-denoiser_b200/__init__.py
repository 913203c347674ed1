"""B200-native normal-guided point-cloud denoising hot path (drop-in for the reference's
Pointcloud/Modules on that path).  Host code is Python/PyTorch; every heavy operation is a hand-written
sm_100a kernel behind the C ABI of include/ngpd.h (libngpd.so).  No CPU fallback."""
from . import _lib  # noqa: F401
from .Decompositionor import Decomposition, Decompositionor  # noqa: F401
from .Denoiser import Denoiser  # noqa: F401
from .GraphBuilder import Graph, GraphBuilder  # noqa: F401
from .Mesh import Mesh  # noqa: F401
from .Noise import Noise  # noqa: F401
from .Object import Pointcloud  # noqa: F401
from .Processor import Processor  # noqa: F401
from .Selector import Selection, Selector  # noqa: F401
from .Utils import GeneralUtils, TorchUtils  # noqa: F401
