from .field import NeRFNetwork, get_encoder, trunc_exp, freq_embed  # noqa: F401
from .mlp import Network  # noqa: F401
from .rendering import NeRFRenderer, sample_pdf  # noqa: F401
