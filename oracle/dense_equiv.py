"""ORACLE O1 -- TEST INFRASTRUCTURE ONLY.

Independent dense-equivalence identities: a sparse convolution over an
active set equals a dense torch.nn.functional convolution of the zero-filled volume, sampled at the
output sites.  Pins offset enumeration (row-major, last dim fastest), weight layout [K^3, Cin, Cout]
and cross-correlation (no flip) WITHOUT using any restated scn code.  CPU only.
"""
import torch
import torch.nn.functional as F


def densify(coords, feats, nb, dims):
    c = feats.shape[1]
    vol = feats.new_zeros((nb, c, dims[0], dims[1], dims[2]))
    if coords.shape[0]:
        vol[coords[:, 3], :, coords[:, 0], coords[:, 1], coords[:, 2]] = feats
    return vol


def sample(vol, coords):
    return vol[coords[:, 3], :, coords[:, 0], coords[:, 1], coords[:, 2]]


def submanifold_conv(coords, feats, weight, nb, dims):
    """weight [27, Cin, Cout] -> rows of the output at the input sites."""
    k3, cin, cout = weight.shape
    w = weight.permute(2, 1, 0).reshape(cout, cin, 3, 3, 3)
    return sample(F.conv3d(densify(coords, feats, nb, dims), w, padding=1), coords)


def strided_conv(coords, feats, weight, nb, dims, out_coords):
    """filter 2 stride 2; weight [8, Cin, Cout]; sampled at out_coords (the coarse sites)."""
    k3, cin, cout = weight.shape
    w = weight.permute(2, 1, 0).reshape(cout, cin, 2, 2, 2)
    return sample(F.conv3d(densify(coords, feats, nb, dims), w, stride=2), out_coords)


def strided_deconv(coarse_coords, coarse_feats, weight, nb, coarse_dims, fine_coords):
    """scn.Deconvolution filter 2 stride 2: fine[p] = coarse[p>>1] @ W[p&1...]; weight [8, Cin, Cout]."""
    k3, cin, cout = weight.shape
    w = weight.permute(1, 2, 0).reshape(cin, cout, 2, 2, 2)
    vol = F.conv_transpose3d(densify(coarse_coords, coarse_feats, nb, coarse_dims), w, stride=2)
    return sample(vol, fine_coords)


def unpool(coarse_coords, coarse_feats, nb, coarse_dims, fine_coords):
    vol = densify(coarse_coords, coarse_feats, nb, coarse_dims)
    vol = vol.repeat_interleave(2, 2).repeat_interleave(2, 3).repeat_interleave(2, 4)
    return sample(vol, fine_coords)
