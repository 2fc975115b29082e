from .encoder import GridEncoder, grid_encode, level_scales
