"""Test-only CPU executor of the FOLDED layer list (dad_3dheads_b200.encoder.fold_state_dict) following the same graph as
csrc/encoder.cu.  Two uses: (1) on CPU it pins the folding (BN fold, K layouts, block-diagonal heads, fusion scalars)
against the oracle; (2) on the GPU box it gives per-layer expected activations, by the library's layer names, so a
mismatch is localised to one kernel."""
import os
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

STAGE_UNITS = (3, 4, 6, 3)


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bf16 and back: what storing an activation / weight as ONE bf16 piece does (encoder mode "bf16")."""
    return t.to(torch.bfloat16).to(t.dtype)


def run_folded(x: torch.Tensor, layers, fusion_w: np.ndarray, dtype=torch.float64, quant=None) -> Dict[str, torch.Tensor]:
    """x [B,3,256,256] -> dict layer-name -> NCHW activation (true channel counts), plus 'params', 'landmarks', 'heat'.

    ``quant`` (e.g. :func:`bf16_round`) emulates a single-piece operand mode: it is applied to the image, to every folded weight
    and to every activation the engine stores as 16-bit pieces (conv outputs after bias / residual / gate / ReLU, the fused
    BiFPN sums, the FusionLayer concat, the pooled features); biases, accumulation and the fp32-kept outputs (heat-map, MLP
    logits) stay unrounded -- the bf16-emulating oracle of BASELINE configs[2]."""
    Q = quant if quant is not None else (lambda t: t)
    L = {n: (Q(torch.from_numpy(w).to(dtype)).permute(0, 3, 1, 2).contiguous(), torch.from_numpy(b).to(dtype))
         for n, w, b in layers}
    acts: Dict[str, torch.Tensor] = {}
    F32_OUT = ("heat", "mlp2")

    def conv(name, t, stride=1, pad=0, relu=False, res=None, mul=None):
        w, b = L[name]
        y = F.conv2d(t, w, b, stride, pad)
        if res is not None:
            y = y + res
        if mul is not None:
            y = y * mul
        if relu:
            y = F.relu(y)
        if name not in F32_OUT:
            y = Q(y)
        acts[name] = y
        return y

    x = Q(x.to(dtype))
    y = conv("stem", x, 2, 3, True)
    acts["stem_conv"] = y
    y = F.max_pool2d(y, 3, 2, 1)
    acts["stem"] = y

    def stage(si, t):
        for ui in range(STAGE_UNITS[si]):
            p = f"s{si + 1}u{ui + 1}"
            stride = 2 if (ui == 0 and si != 0) else 1
            u = conv(p + "c1", t, stride, 0, True)
            u = conv(p + "c2", u, 1, 1, True)
            if ui == 0:          # projection shortcut K-concatenated: [W3 | Wid] [u ; t_strided]
                t = conv(p + "c3", torch.cat([u, t[:, :, ::stride, ::stride]], 1), 1, 0, True)
            else:
                t = conv(p + "c3", u, 1, 0, True, res=t)
        return t

    c2 = stage(0, y)
    c3 = stage(1, c2)
    c4 = stage(2, c3)
    # P3's lateral is composed into b0_p3td by fold_state_dict (no "lat3" layer): the first block's P3 input is c2 itself
    feat = [conv("lat3", c2) if "lat3" in L else c2, conv("lat4", c3), conv("lat5", c4), conv("lat6", c4, 2, 1)]
    feat.append(conv("lat7", feat[3], 2, 1, True))
    near = lambda t, ref: F.interpolate(t, size=ref.shape[2:])
    for li in range(2):
        w1 = fusion_w[li, :8].reshape(2, 4).astype(np.float64)
        w2 = fusion_w[li, 8:].reshape(3, 4).astype(np.float64)
        p3x, p4x, p5x, p6x, p7x = feat
        p = f"b{li}_"
        p7td = p7x
        # top-down nodes: fusion weights are folded into the two weight sets; the low-res product is stored nearest-up-sampled
        # (that is the activation the engine keeps under the "_u" name) and added
        def up(name, low, ref):
            u = near(conv(name, low), ref)          # conv() recorded the half-resolution product under `name`
            if ref.shape[2] < 32 or os.environ.get("DAD3D_TD_PARITY") != "1":
                acts[name] = u                      # the engine stores it up-sampled (4x store); in the opt-in parity mode
            return u                                # large maps keep it at half resolution
        p6td = conv(p + "p6td", p6x, relu=True, res=up(p + "p6td_u", p7td, p6x))
        p5td = conv(p + "p5td", p5x, relu=True, res=up(p + "p5td_u", p6td, p5x))
        p4td = conv(p + "p4td", p4x, relu=True, res=up(p + "p4td_u", p5td, p4x))
        p3td = conv(p + "p3td", p3x, relu=True, res=up(p + "p3td_u", p4td, p3x))
        p4o = conv(p + "p4out", Q(w2[0, 0] * p4x + w2[1, 0] * p4td + w2[2, 0] * near(p3td, p4x)), relu=True)
        p5o = conv(p + "p5out", Q(w2[0, 1] * p5x + w2[1, 1] * p5td + w2[2, 1] * near(p4o, p5x)), relu=True)
        p6o = conv(p + "p6out", Q(w2[0, 2] * p6x + w2[1, 2] * p6td + w2[2, 2] * near(p5o, p6x)), relu=True)
        p7o = conv(p + "p7out", Q(w2[0, 3] * p7x + w2[1, 3] * p7td + w2[2, 3] * near(p6o, p7x)), relu=True)
        feat = [p3td, p4o, p5o, p6o, p7o]
    heat = conv("heat", feat[0], 1, 1)
    hm = F.interpolate(heat, size=c4.shape[2:], mode="bilinear", align_corners=True).sigmoid()
    pad = torch.zeros(hm.shape[0], 128 - hm.shape[1], *hm.shape[2:], dtype=dtype)
    cat = torch.cat([c4, Q(hm), pad, feat[2]], 1)
    acts["cat"] = cat
    f = conv("fusion", cat, mul=c4)
    s4 = stage(3, f)
    gap = Q(F.adaptive_avg_pool2d(s4, 1))
    acts["gap"] = gap
    h = conv("mlp1", gap, relu=True)
    o = conv("mlp2", h).flatten(1)
    acts["params"] = torch.cat([torch.tanh(o[:, :403]) * 3.0, o[:, 403:413]], 1)
    acts["landmarks"] = F.relu(o[:, 413:549]).reshape(-1, 68, 2)
    return acts
