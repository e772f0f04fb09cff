"""KPConv blocks (reference: geotransformer/modules/kpconv/{kpconv,modules,functional,kernel_points}.py)."""
import math

import numpy as np
import torch
import torch.nn as nn

from .. import ops

# k_015_center_3D disposition (15 kernel points, one fixed at the centre, unit sphere): the data table the
# reference ships as geotransformer/modules/kpconv/dispositions/k_015_center_3D.ply, as float32
# (kernel_points.py:423-424 casts to float32 on load).  Pretrained checkpoints carry their own
# `kernel_points` buffers; this table only matters for freshly initialised models.
_K015_CENTER_3D = np.array([
    [0.0, 0.0, 0.0],
    [-0.49820613861083984, 0.41826796531677246, 0.11736718565225601],
    [-0.2412356436252594, -0.3421404957771301, -0.5115481019020081],
    [-0.2828808128833771, -0.5861426591873169, 0.1155322790145874],
    [0.2905403673648834, -0.10093209147453308, -0.5850909948348999],
    [0.4282003939151764, 0.3992988169193268, -0.3068181276321411],
    [-0.6358649134635925, -0.081964410841465, -0.16090403497219086],
    [-0.43181082606315613, -0.14729416370391846, 0.4783095717430115],
    [-0.04466600343585014, 0.2797321379184723, 0.5972330570220947],
    [0.2255241721868515, -0.34462544322013855, 0.5079466104507446],
    [0.6388921141624451, -0.16914905607700348, -0.01190108153969049],
    [-0.2255241423845291, 0.34462544322013855, -0.5079466104507446],
    [0.49054664373397827, 0.2688070237636566, 0.35219207406044006],
    [0.25233083963394165, -0.5970665216445923, -0.12951141595840454],
    [0.03415393829345703, 0.658583402633667, 0.04513958469033241],
], dtype=np.float32)


def load_kernels(radius, num_kpoints, dimension=3, fixed="center"):
    """kernel_points.py:389-455 for the one disposition the model uses: random rotation about z
    (np.random), N(0, 0.01) jitter, scale by `radius`, rotate.  Same numpy RNG consumption as the reference."""
    if num_kpoints != 15 or dimension != 3 or fixed != "center":
        raise NotImplementedError("only the k_015_center_3D kernel disposition is built in")
    kernel_points = _K015_CENTER_3D.copy()
    theta = np.random.rand() * 2 * np.pi
    c, s = np.cos(theta), np.sin(theta)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)
    kernel_points = kernel_points + np.random.normal(scale=0.01, size=kernel_points.shape)
    kernel_points = radius * kernel_points
    kernel_points = np.matmul(kernel_points, R)
    return kernel_points.astype(np.float32)


def maxpool(x, neighbor_indices):
    return ops.maxpool(x, neighbor_indices)


def nearest_upsample(x, upsample_indices):
    return ops.nearest_upsample(x, upsample_indices)


class KPConv(nn.Module):
    """kpconv.py:10-133 (rigid KPConv)."""

    def __init__(self, in_channels, out_channels, kernel_size, radius, sigma, bias=False, dimension=3, inf=1e6, eps=1e-9):
        super().__init__()
        self.kernel_size, self.in_channels, self.out_channels = kernel_size, in_channels, out_channels
        self.radius, self.sigma, self.dimension, self.inf, self.eps = radius, sigma, dimension, inf, eps
        self.weights = nn.Parameter(torch.zeros(kernel_size, in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)
        nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weights)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)
        self.register_buffer("kernel_points", torch.from_numpy(load_kernels(radius, kernel_size, dimension, "center")).float())

    @torch.no_grad()
    def forward(self, s_feats, q_points, s_points, neighbor_indices):
        return ops.kpconv(s_feats, q_points, s_points, neighbor_indices, self.weights, self.bias, self.kernel_points, self.sigma)


class GroupNorm(nn.Module):
    """modules.py:33-50: nn.GroupNorm over (1, C, N), i.e. statistics over all stacked rows."""

    def __init__(self, num_groups, num_channels):
        super().__init__()
        self.num_groups, self.num_channels = num_groups, num_channels
        self.norm = nn.GroupNorm(num_groups, num_channels)  # parameter container (same state_dict keys)

    @torch.no_grad()
    def forward(self, x, add=None, act=None):
        return ops.group_norm(x, self.num_groups, self.norm.weight, self.norm.bias, self.norm.eps, add=add, act=act)


class UnaryBlock(nn.Module):
    """modules.py:53-83."""

    def __init__(self, in_channels, out_channels, group_norm, has_relu=True, bias=True, layer_norm=False):
        super().__init__()
        if layer_norm:
            raise NotImplementedError("layer_norm=True is not used by the GaussReg model")
        self.in_channels, self.out_channels, self.group_norm = in_channels, out_channels, group_norm
        self.mlp = nn.Linear(in_channels, out_channels, bias=bias)
        self.norm = GroupNorm(group_norm, out_channels)
        self.leaky_relu = nn.LeakyReLU(0.1) if has_relu else None

    @torch.no_grad()
    def forward(self, x, add=None, act_after_add=None):
        # Linear -> GroupNorm (+ add) -> activation in one C-ABI call; the norm's statistics ride in the product's epilogue
        return ops.unary_block(self, x, add=add, act_after_add=act_after_add)


class LastUnaryBlock(nn.Module):
    """modules.py:86-101."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.mlp = nn.Linear(in_channels, out_channels, bias=bias)

    @torch.no_grad()
    def forward(self, x):
        return ops.linear(x, self.mlp.weight, self.mlp.bias)


class ConvBlock(nn.Module):
    """modules.py:104-146."""

    def __init__(self, in_channels, out_channels, kernel_size, radius, sigma, group_norm, negative_slope=0.1, bias=True,
                 layer_norm=False):
        super().__init__()
        if layer_norm or negative_slope != 0.1:
            raise NotImplementedError
        self.in_channels, self.out_channels = in_channels, out_channels
        self.KPConv = KPConv(in_channels, out_channels, kernel_size, radius, sigma, bias=bias)
        self.norm = GroupNorm(group_norm, out_channels)
        self.leaky_relu = nn.LeakyReLU(negative_slope=negative_slope)

    @torch.no_grad()
    def forward(self, s_feats, q_points, s_points, neighbor_indices):
        return ops.kpconv_block(self.KPConv, self.norm, s_feats, q_points, s_points, neighbor_indices)


class ResidualBlock(nn.Module):
    """modules.py:149-225 (bottleneck: unary1 -> KPConv -> GN+LReLU -> unary2, + shortcut, LReLU)."""

    def __init__(self, in_channels, out_channels, kernel_size, radius, sigma, group_norm, strided=False, bias=True,
                 layer_norm=False):
        super().__init__()
        if layer_norm:
            raise NotImplementedError
        self.in_channels, self.out_channels, self.strided = in_channels, out_channels, strided
        mid = out_channels // 4
        self.unary1 = UnaryBlock(in_channels, mid, group_norm, bias=bias) if in_channels != mid else nn.Identity()
        self.KPConv = KPConv(mid, mid, kernel_size, radius, sigma, bias=bias)
        self.norm_conv = GroupNorm(group_norm, mid)
        self.unary2 = UnaryBlock(mid, out_channels, group_norm, has_relu=False, bias=bias)
        if in_channels != out_channels:
            self.unary_shortcut = UnaryBlock(in_channels, out_channels, group_norm, has_relu=False, bias=bias)
        else:
            self.unary_shortcut = nn.Identity()
        self.leaky_relu = nn.LeakyReLU(0.1)

    @torch.no_grad()
    def forward(self, s_feats, q_points, s_points, neighbor_indices):
        x = self.unary1(s_feats)
        x = ops.kpconv_block(self.KPConv, self.norm_conv, x, q_points, s_points, neighbor_indices)
        shortcut = maxpool(s_feats, neighbor_indices) if self.strided else s_feats
        if not isinstance(self.unary_shortcut, nn.Identity):
            shortcut = self.unary_shortcut(shortcut)
        # LeakyReLU(GN(Linear(x)) + shortcut): the add and the activation ride in the GroupNorm apply pass
        return self.unary2(x, add=shortcut, act_after_add="leaky_relu")
