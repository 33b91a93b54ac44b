"""CPU check of the index arithmetic behind csrc/convflat.cu (the flattened, zero-separated implicit GEMM of the small k4 s2
layers: networks/networks.py:157-165): the operand layout of flat_pack_kernel, the shift table, and the weight packers of
ops_conv are restated in torch and the resulting GEMM is compared with torch's own convolutions.  No GPU, no kernel launch."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from genre_shapehd_b200 import _lib, ops_conv


def positions(b, d, h, w):
    lead = ctypes.c_int(0)
    p = _lib.load().genre_b200_convflat_positions(b, d, h, w, ctypes.byref(lead))
    return int(p), lead.value


def operand_of(sources, d, h, w, subvol):
    """torch restatement of flat_pack_kernel: [channels][P] fp64 (one part), separators / lead / tail zero"""
    b = sources[0].shape[0]
    npos, lead = positions(b, d, h, w)
    chans = []
    for x in sources:
        if subvol:     # 8 parity sub-volumes as channel blocks, sub-volume s = (pz*2+py)*2+px
            x = torch.cat([x[:, :, pz::2, py::2, px::2] for pz in (0, 1) for py in (0, 1) for px in (0, 1)], dim=1)
        chans.append(x)
    x = torch.cat(chans, dim=1).double()
    c = x.shape[1]
    padded = F.pad(x, (0, 1, 0, 1, 0, 1))                       # one zero separator after every row, plane and volume
    flat = padded.permute(1, 0, 2, 3, 4).reshape(c, -1)
    act = torch.zeros(c + (-c) % 16, npos, dtype=torch.float64)
    act[:c, lead:lead + flat.shape[1]] = flat
    return act, lead


def run_flat(act, lead, wpack, b, d, h, w, cout, npad, transposed):
    """the GEMM of convflat_kernel on the restated operand; returns NCDHW fp64"""
    shifts = ops_conv.flat_shifts(h, w, transposed)
    m = b * (d + 1) * (h + 1) * (w + 1)
    classes, ntiles, ksteps = wpack.shape[0], wpack.shape[1], wpack.shape[2]
    assert ksteps * 16 == act.shape[0]
    od, oh, ow = (2 * d, 2 * h, 2 * w) if transposed else (d, h, w)
    out = torch.zeros(b, cout, od, oh, ow, dtype=torch.float64)
    q = torch.arange(m)
    x, y, z, bb = q % (w + 1), (q // (w + 1)) % (h + 1), (q // ((w + 1) * (h + 1))) % (d + 1), q // ((w + 1) * (h + 1) * (d + 1))
    valid = (x < w) & (y < h) & (z < d)
    for cls in range(classes):
        for nt in range(ntiles):
            acc = torch.zeros(m, npad, dtype=torch.float64)
            for ks in range(ksteps):
                grp = cls if transposed else (ks * 2) // (act.shape[0] // 8 // 8) % 8
                for tap in range(8):
                    wt = wpack[cls, nt, ks, tap].double()                       # [kk, ng, r, e]
                    wmat = wt.permute(0, 3, 1, 2).reshape(16, npad)             # [k = kk*8 + e][n = ng*8 + r]
                    sh = shifts[grp][tap]
                    a = act[ks * 16:(ks + 1) * 16, lead + sh:lead + sh + m]     # [16, m]
                    acc += a.t() @ wmat
            pz, py, px = ((cls >> 2) & 1, (cls >> 1) & 1, cls & 1) if transposed else (0, 0, 0)
            s = 2 if transposed else 1
            n0 = nt * npad
            nn = min(npad, cout - n0)
            out[bb[valid], n0:n0 + nn, s * z[valid] + pz, s * y[valid] + py, s * x[valid] + px] = acc[valid][:, :nn]
    return out


@pytest.mark.parametrize("dims,cins,cout", [((4, 4, 4), (16, 16), 24), ((2, 3, 5), (8,), 70), ((8, 8, 8), (24, 8), 90), ((1, 1, 1), (16,), 8)])
def test_flat_transposed_conv_arithmetic(dims, cins, cout):
    torch.manual_seed(0)
    b = 3
    d, h, w = dims
    sources = [torch.randn(b, c, d, h, w) for c in cins]
    weight = torch.randn(sum(cins), cout, 4, 4, 4)
    npad = ops_conv.flat_npad(cout)
    wpack = ops_conv.pack_flat_convt_weights(weight, npad, 4)      # group 4: keep the fp32 values
    act, lead = operand_of(sources, d, h, w, False)
    got = run_flat(act, lead, wpack, b, d, h, w, cout, npad, True)
    ref = F.conv_transpose3d(torch.cat(sources, 1).double(), weight.double(), stride=2, padding=1)
    assert got.shape == ref.shape
    assert (got - ref).abs().max() < 1e-9


@pytest.mark.parametrize("dims,cin,cout", [((8, 8, 8), 16, 24), ((4, 6, 2), 32, 100), ((16, 16, 16), 16, 64), ((2, 2, 2), 16, 8)])
def test_flat_strided_conv_arithmetic(dims, cin, cout):
    torch.manual_seed(1)
    b = 2
    x = torch.randn(b, cin, *dims)
    d, h, w = (v // 2 for v in dims)
    weight = torch.randn(cout, cin, 4, 4, 4)
    npad = ops_conv.flat_npad(cout)
    wpack = ops_conv.pack_flat_conv_weights(weight, npad, 4)
    act, lead = operand_of([x], d, h, w, True)
    got = run_flat(act, lead, wpack, b, d, h, w, cout, npad, False)
    ref = F.conv3d(x.double(), weight.double(), stride=2, padding=1)
    assert got.shape == ref.shape
    assert (got - ref).abs().max() < 1e-9


def test_flat_operand_bounds():
    """every shifted read of every CTA stays inside the operand arrays (lead >= halo, tail >= halo + tile round-up)"""
    for (b, d, h, w) in [(16, 8, 8, 8), (16, 4, 4, 4), (1, 1, 1, 1), (3, 2, 7, 5), (16, 16, 16, 16)]:
        npos, lead = positions(b, d, h, w)
        halo = (h + 1) * (w + 1) + (w + 1) + 1
        m = b * (d + 1) * (h + 1) * (w + 1)
        assert lead >= halo and lead % 8 == 0
        assert max(abs(s) for row in ops_conv.flat_shifts(h, w, True) + ops_conv.flat_shifts(h, w, False) for s in row) <= halo
        for mt in (1, 2):
            tiles = -(-m // (128 * mt))
            assert lead + tiles * 128 * mt + halo <= npos
