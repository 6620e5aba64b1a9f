"""Packs the decode-path tensors of a STRIVE `TrafficModel.state_dict()` into the kernel layout.

Segment order = `enum Seg` in csrc/common.cuh (checked at load time against strive_model_layout()).
`_T` segments are [in][out] (transposed nn.Linear weights: coalesced forward GEMVs), `_N` segments are the
native [out][in] matrices (or column slices) used by the data-gradient GEMVs of the backward pass.
Reference key names: src/models/traffic_model.py:69-156, src/models/interaction_net.py:29-119, models/common.py:8-39.
"""
import torch


def _pad_rows(t, rows):
    if t.size(0) == rows:
        return t
    out = torch.zeros((rows, t.size(1)), dtype=t.dtype, device=t.device)
    out[:t.size(0)] = t
    return out


def _r4(x):
    return (x + 3) & ~3


def decode_segments(sd, NC):
    """List of (name, 2-D/1-D float32 CPU tensor) in enum Seg order."""
    g = lambda k: sd[k].detach().to(torch.float32).cpu()
    segs = []
    for l in range(6):
        w = g('map_conv.%d.weight' % (3 * l))                       # (Cout,Cin,k,k)
        segs.append(('CW%d' % l, w.permute(1, 2, 3, 0).reshape(-1, w.size(0))))
        segs.append(('CB%d' % l, g('map_conv.%d.bias' % (3 * l))))
        segs.append(('GG%d' % l, g('map_conv.%d.weight' % (3 * l + 1))))
        segs.append(('GB%d' % l, g('map_conv.%d.bias' % (3 * l + 1))))
    segs.append(('FCW', g('map_feature.weight').t()))
    segs.append(('FCB', g('map_feature.bias')))
    p = 'decoder_net.mlp_in.net.'
    w0 = g(p + '0.weight')                                           # (128, 64+64+NC+32+2)
    in0_rows = _r4(64 + 64 + NC + 32 + 2)
    assert w0.size(1) == 64 + 64 + NC + 32 + 2, 'decoder_net.mlp_in input width does not match num_classes'
    segs += [('IN0_T', _pad_rows(w0.t(), in0_rows)), ('IN0_B', g(p + '0.bias')),
             ('IN_LN1_G', g(p + '1.weight')), ('IN_LN1_B', g(p + '1.bias')),
             ('IN3_T', g(p + '3.weight').t()), ('IN3_B', g(p + '3.bias')),
             ('IN_LN4_G', g(p + '4.weight')), ('IN_LN4_B', g(p + '4.bias')),
             ('IN6_T', g(p + '6.weight').t()), ('IN6_B', g(p + '6.bias')),
             ('IN0_N_PF', w0[:, 0:64]), ('IN0_N_Z', w0[:, 128 + NC:128 + NC + 32]),
             ('IN3_N', g(p + '3.weight')), ('IN6_N', g(p + '6.weight'))]
    p = 'decoder_net.msg.0.edge_mlp.net.'
    w1 = g(p + '0.weight')                                           # (128, 2*(64+NC)+4)
    assert w1.size(1) == 2 * (64 + NC) + 4
    segs += [('E0_T_XI', w1[:, 0:64].t()), ('E0_T_XJ', w1[:, 64:128].t()),
             ('E0_T_SEMI', w1[:, 128:128 + NC].t()), ('E0_T_SEMJ', w1[:, 128 + NC:128 + 2 * NC].t()),
             ('E0_T_REL', w1[:, 128 + 2 * NC:].t()), ('E0_B', g(p + '0.bias')),
             ('E0_N_XI', w1[:, 0:64]), ('E0_N_XJ', w1[:, 64:128]),
             ('E_LN1_G', g(p + '1.weight')), ('E_LN1_B', g(p + '1.bias')),
             ('E3_T', g(p + '3.weight').t()), ('E3_N', g(p + '3.weight')), ('E3_B', g(p + '3.bias')),
             ('E_LN4_G', g(p + '4.weight')), ('E_LN4_B', g(p + '4.bias')),
             ('E6_T', g(p + '6.weight').t()), ('E6_N', g(p + '6.weight')), ('E6_B', g(p + '6.bias'))]
    p = 'decoder_net.msg.0.update_mlp.net.'
    wu = g(p + '0.weight')                                           # (128, 64+64+NC)
    u0_rows = _r4(64 + 64 + NC)
    segs += [('U0_T', _pad_rows(wu.t(), u0_rows)), ('U0_B', g(p + '0.bias')),
             ('U_LN1_G', g(p + '1.weight')), ('U_LN1_B', g(p + '1.bias')),
             ('U3_T', g(p + '3.weight').t()), ('U3_B', g(p + '3.bias')),
             ('U0_N_X', wu[:, 0:64]), ('U0_N_AGGR', wu[:, 64:128]), ('U3_N', g(p + '3.weight'))]
    p = 'decoder_net.mlp_out.net.'
    segs += [('O0_T', g(p + '0.weight').t()), ('O0_B', g(p + '0.bias')),
             ('O_LN1_G', g(p + '1.weight')), ('O_LN1_B', g(p + '1.bias')),
             ('O3_T', g(p + '3.weight').t()), ('O3_B', g(p + '3.bias')),
             ('O_LN4_G', g(p + '4.weight')), ('O_LN4_B', g(p + '4.bias')),
             ('O6_N', g(p + '6.weight')), ('O6_B', g(p + '6.bias')),
             ('O0_N', g(p + '0.weight')), ('O3_N', g(p + '3.weight'))]
    for l in range(3):
        wi, wh = g('decoder_memory.weight_ih_l%d' % l), g('decoder_memory.weight_hh_l%d' % l)
        segs += [('GI_T%d' % l, wi.t()), ('GH_T%d' % l, wh.t()),
                 ('GBI%d' % l, g('decoder_memory.bias_ih_l%d' % l)), ('GBH%d' % l, g('decoder_memory.bias_hh_l%d' % l)),
                 ('GI_N%d' % l, wi), ('GH_N%d' % l, wh)]
    return segs


def pack_decode_weights(sd, NC):
    """-> (blob float32 CPU 1-D, [segment sizes])  every segment starts on a 16-byte boundary."""
    segs = decode_segments(sd, NC)
    sizes = [int(t.numel()) for _, t in segs]
    total = sum(_r4(s) for s in sizes)
    blob = torch.zeros(total, dtype=torch.float32)
    off = 0
    for (_, t), s in zip(segs, sizes):
        blob[off:off + s] = t.contiguous().reshape(-1)
        off += _r4(s)
    return blob, sizes


def _split(w):
    hi = w.to(torch.bfloat16)
    lo = (w - hi.to(torch.float32)).to(torch.bfloat16)
    return hi, lo


def _bytes(t):
    return t.contiguous().view(torch.int16).reshape(-1).view(torch.uint8)


CONV1_F = 127 * 65536      # fixed-point scale of the conv1 weights: the top digit of max|w| is exactly 127


def conv1_fixed_point(w):
    """(16,4,7,7) float32 -> (digits int64 (3,16,4,7,7) in [-128,127], scale float32 (16)):  w ~= scale[o] * sum_d digits[d] 256^d,
    q = round(w / max|w[o]| * 127 * 2^16) in balanced base-256 digits, |w - scale q| <= 2^-24 max|w[o]| (+ the fp32 rounding of
    scale = max|w[o]| / (127 * 2^16), a common factor of the channel).  The sums over the 196 binary inputs of one digit plane
    stay below 2^15, so the kernel epilogue recombines the planes exactly in fp32 arithmetic."""
    w64 = w.to(torch.float64)
    s = w64.abs().reshape(w.size(0), -1).max(dim=1).values
    s = torch.where(s > 0, s, torch.ones_like(s))
    q = torch.round(w64 / s.view(-1, 1, 1, 1) * float(CONV1_F)).to(torch.int64)
    digits = []
    for i in range(3):
        d = ((q + 128) % 256) - 128 if i < 2 else q
        digits.append(d)
        q = (q - d) // 256
    assert int(q.abs().max()) == 0 and int(digits[2].abs().max()) <= 127
    return torch.stack(digits), (s / float(CONV1_F)).to(torch.float32)


def pack_conv1_fixed_point(w):
    """conv1 operand B of tcgen05.mma kind::i8: [ky 7][khalf 2][n = 16 d + o][k = 4 kx_l + c] int8, taps kx = 4 khalf + kx_l
    (kx = 7 is a zero pad), followed by the 16 fp32 scales."""
    digits, scale = conv1_fixed_point(w)                                        # (3,16,4,7,7)
    dg = torch.cat([digits, torch.zeros(3, 16, 4, 7, 1, dtype=torch.int64)], dim=4)   # pad kx to 8
    t = dg.reshape(3, 16, 4, 7, 2, 4)                                           # (d, o, c, ky, khalf, kxl)
    t = t.permute(3, 4, 0, 1, 5, 2).contiguous()                                # (ky, khalf, d, o, kxl, c)
    b = t.to(torch.int8).reshape(-1).view(torch.uint8)
    return torch.cat([b, scale.contiguous().view(torch.uint8).reshape(-1)])


def pack_tc_weights(sd):
    """Map-encoder weights in the tcgen05 K-major no-swizzle operand layout (csrc/mapenc_tc.cu).
    conv1: int8 digit planes of a 31-bit fixed-point representation (pack_conv1_fixed_point).
    conv2..fc: bf16 hi/lo split, the hi and lo halves STACKED ALONG N (rows n' = prec * NCH + n), so one MMA multiplies an A tile
    with both;  conv2..4: [nchunk][c2][tap][khalf 2][prec 2][ngroup NCH/8][r 8][k 8], n = NCH nchunk + 8 ngroup + r, c = 16 c2 + 8 khalf + k,
    NCH = 32 output channels per CTA for conv2 and 64 for conv3 / conv4."""
    g = lambda k: sd[k].detach().to(torch.float32).cpu()
    out = []
    out.append(pack_conv1_fixed_point(g('map_conv.0.weight')))
    for li, ks, nch in ((1, 5, 32), (2, 5, 64), (3, 3, 64)):
        w = g('map_conv.%d.weight' % (3 * li))                              # (Cout, Cin, ks, ks)
        cout, cin = w.size(0), w.size(1)
        parts = []
        for p in _split(w):
            t = p.reshape(cout // nch, nch // 8, 8, cin // 16, 2, 8, ks * ks)  # (nchunk, ngroup, r, c2, khalf, k, tap)
            parts.append(t.permute(0, 3, 6, 4, 1, 2, 5))                    # (nchunk, c2, tap, khalf, ngroup, r, k)
        t = torch.stack(parts, dim=4)                                       # (nchunk, c2, tap, khalf, prec, ngroup, r, k)
        out.append(_bytes(t))
    # conv5 / conv6 / fc as GEMMs with K = tap * Cin + c in 64-wide chunks: [kchunk][prec][kg 8][ng Cout/8][r 8][kk 8]
    def gemm_pack(wk):                                                      # wk (Cout, K)
        cout, K = wk.shape
        parts = []
        for p in _split(wk):
            t = p.reshape(cout // 8, 8, K // 64, 8, 8)                      # (ng, r, kchunk, kg, kk)
            parts.append(t.permute(2, 3, 0, 1, 4))                          # (kchunk, kg, ng, r, kk)
        return _bytes(torch.stack(parts, dim=1))                            # (kchunk, prec, kg, ng, r, kk)
    for li in (4, 5):
        w = g('map_conv.%d.weight' % (3 * li))                              # (Cout, Cin, 3, 3) -> k = (ky*3+kx)*Cin + c
        out.append(gemm_pack(w.permute(0, 2, 3, 1).reshape(w.size(0), -1)))
    wf = g('map_feature.weight').reshape(64, 128, 2, 2)                     # flatten index of the reference = c*4 + y*2 + x
    out.append(gemm_pack(wf.permute(0, 2, 3, 1).reshape(64, -1)))           # NHWC order: k = (y*2+x)*128 + c
    # conv3 once more for the CTA-pair kernel (tcgen05 cta_group::2: each CTA of a pair stages HALF of every B operand):
    # [rank 2][c2 2][tap 25][khalf 2][64 rows][8 k] with rows = W_hi rows 32 rank .. 32 rank + 31, then W_lo rows 32 rank .. 32 rank + 31
    w = g('map_conv.6.weight')
    hi, lo = [p.reshape(8, 8, 2, 2, 8, 25).permute(2, 5, 3, 0, 1, 4) for p in _split(w)]     # (c2, tap, khalf, ngroup, r, k)
    out.append(_bytes(torch.stack([torch.cat([hi[:, :, :, 4 * rank:4 * rank + 4], lo[:, :, :, 4 * rank:4 * rank + 4]], dim=3) for rank in (0, 1)])))
    return torch.cat(out)
