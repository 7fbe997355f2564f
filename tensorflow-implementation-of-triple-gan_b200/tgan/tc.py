"""bf16 tensor-core (tcgen05 / TMEM / TMA) routing of the dense contractions.

`ops.conv2d` / `ops.conv2d_transpose` call in here when ctx.math == 'bf16'.  A layer is eligible when the
implicit-GEMM kernel's tiling covers it (stride 1, 128-pixel tiles that align with image rows, >= 16
output channels); everything else (Cout in {1,3,10}, stride-2 convs) stays on the fp32 SIMT path.
"""
from .core import ctx  # noqa: F401


def conv_eligible(geom, x):
    return False


def deconv_eligible(geom, x):
    return False


def conv_fwd(x, w, geom):
    raise NotImplementedError


def conv_bwd(x, w, geom, dz):
    raise NotImplementedError


def deconv_fwd(x, w, geom):
    raise NotImplementedError


def deconv_bwd(x, w, geom, dy):
    raise NotImplementedError
