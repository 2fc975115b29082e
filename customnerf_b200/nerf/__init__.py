from .field import NeRFNetwork, get_encoder, trunc_exp, freq_embed  # noqa: F401
from .mlp import Network  # noqa: F401
from .rendering import NeRFRenderer, sample_pdf  # noqa: F401
from .provider_utils import get_rays  # noqa: F401
