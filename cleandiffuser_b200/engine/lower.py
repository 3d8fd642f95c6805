"""Lower a denoiser ``nn.Module`` to the fused-operator program libcds executes.

The module tree (same names/shapes as the reference's, so user checkpoints load) is walked once per
(model, batch, option set); the result is a ``Program``:

* ``ops``        -- ``cabi.Op`` list for ONE reverse iteration (denoiser + solver update)
* ``packers``    -- re-run when parameter versions change (training mutates weights between calls):
                    copy/transpose parameters into the K-major layouts the kernels read
* ``per_call``   -- run at every ``sample()``: fill the per-iteration tables (time-conditioning rows evaluated
                    with the model's OWN ``map_noise`` / ``map_emb`` / ``emb_mlp`` modules on one row per
                    iteration -- inside ``sample()`` t is batch-constant, diffusionsde.py:528/:874 -- and the
                    per-trajectory condition terms)
* ``keep``       -- every device tensor the ops point into (workspace, packed weights, tables)

What is hoisted out of the iteration loop:
  JannerUNet1d w/o condition : the whole map_noise -> map_emb -> emb_mlp chain is a [n_iters, sum C_out] table.
  ChiUNet1d                  : cond_encoder = Linear(Mish(cat[time, obs])) splits exactly into a per-iteration
                               row plus a per-trajectory, iteration-invariant row (Mish is elementwise).
  DQLMlp                     : the first Linear over cat[x, time_mlp(t), obs] splits the same way.
  DiT1d / Janner + condition : map_emb(map_noise(t) + cond) is nonlinear in the sum -> small per-trajectory
                               GEMMs stay inside the iteration (0.5 % of the FLOPs).
"""
import ctypes as C
from typing import Callable, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import cabi


class Unsupported(Exception):
    """The module / option set is outside what the kernels cover -> caller uses the PyTorch path."""


class View:
    """A channels-last activation: element (b, l, c) at base + b*bstride + l*lstride + c (elements of its dtype)."""

    def __init__(self, tensor: torch.Tensor, L: int, Cn: int, offset: int = 0, bstride: Optional[int] = None,
                 lstride: Optional[int] = None, tf32: bool = False):
        assert tensor.dtype in (torch.float32, torch.bfloat16)
        # tf32: fp32 storage that feeds a TF32 tensor-core operator -> producers round to TF32 when they write it (cds_dtype
        # CDS_TF32); readers treat it as fp32
        self.tf32 = bool(tf32) and tensor.dtype == torch.float32
        # module dims may be numpy ints (np.cumprod(dim_mult)); ctypes wants python ints
        self.t, self.L, self.C, self.offset = tensor, int(L), int(Cn), int(offset)
        self.lstride = self.C if lstride is None else int(lstride)
        self.bstride = self.L * self.lstride if bstride is None else int(bstride)

    @property
    def ptr(self):
        return self.t.data_ptr() + self.t.element_size() * self.offset

    @property
    def dtype(self):
        return cabi.BF16 if self.t.dtype == torch.bfloat16 else (cabi.TF32 if self.tf32 else cabi.F32)

    def channels(self, start: int, count: int) -> "View":
        return View(self.t, self.L, count, self.offset + int(start), self.bstride, self.lstride, self.tf32)


def _vec(step: Optional[torch.Tensor] = None, sample: Optional[torch.Tensor] = None, col: int = 0) -> cabi.Vec:
    """step: [n_iters, W] table, sample: [rows, W] table; the vector is columns col.. of both."""
    v = cabi.Vec()
    col = int(col)
    if step is not None:
        v.step, v.step_stride = step.data_ptr() + 4 * col, step.stride(0)
    if sample is not None:
        v.sample, v.sample_stride = sample.data_ptr() + 4 * col, sample.stride(0)
    return v


def _const_vec(t: Optional[torch.Tensor]) -> cabi.Vec:
    v = cabi.Vec()
    if t is not None:
        v.step, v.step_stride = t.data_ptr(), 0
    return v


class WSpec:
    """Where a GEMM weight comes from; packed lazily in the layout of the kernel that ends up running the op.

    kn() -> fp32 [K = taps*C_in, N = C_out*phases]   (CUDA-core kernel: K rows, N contiguous)
    nk() -> [taps, C_out, C_in]                      (tensor-core kernel: K contiguous; cast to bf16 / rounded to TF32), or None"""

    def __init__(self, kn: Callable[[], torch.Tensor], nk: Optional[Callable[[], torch.Tensor]] = None):
        self.kn, self.nk = kn, nk


def _pad_cin(w: torch.Tensor, cin: Optional[int]) -> torch.Tensor:
    """Zero-pad dim 1 (input channels) of a [Cout, Cin, ...] weight up to ``cin`` (x_t is fed 32-channel padded)."""
    if cin is None or cin == w.shape[1]:
        return w
    out = torch.zeros((w.shape[0], cin, *w.shape[2:]), device=w.device, dtype=w.dtype)
    out[:, :w.shape[1]] = w
    return out


def w_conv(conv: nn.Conv1d, cin: Optional[int] = None) -> WSpec:           # [Cout, Cin, k]
    return WSpec(lambda: _pad_cin(conv.weight, cin).permute(2, 1, 0).reshape(-1, conv.weight.shape[0]),
                 lambda: _pad_cin(conv.weight, cin).permute(2, 0, 1))


def w_linear(weight: torch.Tensor, cols: Optional[slice] = None) -> WSpec:    # [out, in]
    pick = (lambda: weight) if cols is None else (lambda: weight[:, cols])
    return WSpec(lambda: pick().t(), lambda: pick()[None])


def w_rows(make_out_in: Callable[[], torch.Tensor]) -> WSpec:                 # callable -> [out, in]
    return WSpec(lambda: make_out_in().t(), lambda: make_out_in()[None])


def w_convT(conv: nn.ConvTranspose1d) -> WSpec:
    """ConvTranspose1d(k=4, s=2, p=1) as a 3-tap conv with two output phases:
    out[2m]   = x[m-1] W[..,3] + x[m] W[..,1]        out[2m+1] = x[m] W[..,2] + x[m+1] W[..,0]"""
    def make():
        w = conv.weight                                      # [Cin, Cout, 4]
        cin, cout, _ = w.shape
        p = torch.zeros(3, cin, 2 * cout, device=w.device, dtype=w.dtype)
        p[0, :, :cout] = w[:, :, 3]
        p[1, :, :cout] = w[:, :, 1]
        p[1, :, cout:] = w[:, :, 2]
        p[2, :, cout:] = w[:, :, 0]
        return p.reshape(3 * cin, 2 * cout)

    def make_nk():                                           # [3][2*Cout][Cin]
        return make().reshape(3, conv.weight.shape[0], -1).permute(0, 2, 1)
    return WSpec(make, make_nk)


def round_tf32(w: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest TF32-representable fp32 (10-bit mantissa, ties away from zero like cvt.rna.tf32.f32).  tcgen05
    kind::tf32 ignores the low 13 mantissa bits of its operands, i.e. truncates; weights are rounded once here so that
    their part of the error is unbiased."""
    bits = w.detach().to(torch.float32).contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


class Program:
    def __init__(self, device: torch.device, rows: int, n_iters: int, math: int = cabi.MATH_FP32):
        self.device, self.rows, self.n_iters, self.math = device, rows, n_iters, math
        self.ops: List[cabi.Op] = []
        self.keep: List[torch.Tensor] = []
        self.packers: List[Callable[[], None]] = []
        self.per_call: List[Callable[[object], None]] = []

    # ---- memory ---------------------------------------------------------------------------
    def buf(self, *shape, dtype=torch.float32) -> torch.Tensor:
        t = torch.zeros(tuple(int(s) for s in shape), device=self.device, dtype=dtype)
        self.keep.append(t)
        return t

    @property
    def act_dtype(self):
        """Intermediate activations: bf16 on CDS_MATH_BF16_TC programs, fp32 otherwise (CUDA-core and TF32 programs)."""
        return torch.bfloat16 if self.math == cabi.MATH_BF16_TC else torch.float32

    @property
    def tc(self) -> bool:
        """Does this program ask for the tensor-core kernels?"""
        return self.math in cabi.TC_MODES

    @property
    def k_align(self) -> int:
        """Channel granularity of a tensor-core operand row (64 bytes): what x_t is padded to."""
        return 16 if self.math == cabi.MATH_TF32_TC else 32

    def act(self, L: int, Cn: int, dtype=None) -> View:
        """A (rows, L, Cn) activation.  Without an explicit dtype it is an inter-layer activation in the program's activation
        type (bf16 / TF32-rounded fp32 / fp32); an explicit torch.float32 is kept exact (predictions, DiT's residual stream)."""
        return View(self.buf(self.rows, L, Cn, dtype=self.act_dtype if dtype is None else dtype), L, Cn,
                    tf32=dtype is None and self.math == cabi.MATH_TF32_TC)

    def packed(self, make: Callable[[], torch.Tensor], dtype=torch.float32) -> torch.Tensor:
        """Persistent packed copy of parameters; ``make`` is re-evaluated into it when weights change."""
        with torch.no_grad():
            t = make().detach().to(device=self.device, dtype=dtype).contiguous().clone()
        self.keep.append(t)

        def refresh():
            with torch.no_grad():
                t.copy_(make())
        self.packers.append(refresh)
        return t

    # ---- operator emitters --------------------------------------------------------------------
    def conv(self, x: View, w: WSpec, out: View, *, taps=1, stride=1, pad=0, phases=1, L_out=None,
             bias: Optional[cabi.Vec] = None, gn: Optional[nn.Module] = None, act=cabi.ACT_NONE,
             scale: Optional[cabi.Vec] = None, shift: Optional[cabi.Vec] = None, res: Optional[View] = None,
             res_conv=None, in_batch_mod=0, res_batch_mod=0, rows=None, sample_row_div=0):
        """Emit one CDS_OP_CONV.  ``res_conv`` = (View, WSpec of the 1x1 weight, bias tensor maker)."""
        op = cabi.Op()
        op.kind = cabi.OP_CONV
        c = op.u.conv
        c.batch = self.rows if rows is None else rows
        c.L_in, c.L_out = x.L, int(out.L // phases if L_out is None else L_out)
        c.C_in, c.C_out = x.C, out.C
        c.taps, c.stride, c.pad, c.phases = int(taps), int(stride), int(pad), int(phases)
        c.in_batch_mod = int(in_batch_mod)
        c.sample_row_div = int(sample_row_div)
        c.in_, c.in_bstride, c.in_lstride, c.in_dtype = x.ptr, x.bstride, x.lstride, x.dtype
        if bias is not None:
            c.bias = bias
        if gn is not None:
            gamma, beta = self.packed(lambda: gn.weight), self.packed(lambda: gn.bias)
            c.groups, c.gn_gamma, c.gn_beta, c.gn_eps = int(gn.num_groups), gamma.data_ptr(), beta.data_ptr(), gn.eps
        c.act = act
        if scale is not None:
            c.scale = scale
        if shift is not None:
            c.shift = shift
        if res is not None:
            c.res, c.res_bstride, c.res_lstride, c.res_dtype = res.ptr, res.bstride, res.lstride, res.dtype
        c.res_batch_mod = int(res_batch_mod)
        if res_conv is not None:
            rx = res_conv[0]
            c.res_in, c.res_in_bstride, c.res_in_lstride, c.res_C = rx.ptr, rx.bstride, rx.lstride, rx.C
            c.res_in_dtype = rx.dtype
            c.res_w = 1          # placeholder so that eligibility sees a shortcut conv; real pointer set below
        c.out, c.out_bstride, c.out_lstride, c.out_dtype = out.ptr, out.bstride, out.lstride, out.dtype

        # kernel family: tensor cores when the program asks for them AND this op qualifies, else CUDA cores
        use_tc = False
        if self.tc and w.nk is not None and (res_conv is None or res_conv[1].nk is not None):
            c.math = self.math
            use_tc = bool(cabi.load().cds_conv_tc_supported(C.byref(c)))
        c.math = self.math if use_tc else cabi.MATH_FP32
        tf32 = self.math == cabi.MATH_TF32_TC
        wdt = torch.float32 if tf32 else torch.bfloat16            # tensor-core weight element
        wfix = round_tf32 if tf32 else (lambda t: t)
        if use_tc:
            wt = self.packed(lambda: wfix(w.nk().reshape(taps * out.C * phases, x.C)), wdt)
        else:
            wt = self.packed(w.kn)
            assert wt.shape == (taps * x.C, out.C * phases), (wt.shape, taps, x.C, out.C, phases)
        c.w = wt.data_ptr()
        if res_conv is not None:
            rw = res_conv[1]
            rwt = self.packed(lambda: wfix(rw.nk().reshape(out.C, rx.C)), wdt) if use_tc else self.packed(rw.kn)
            c.res_w, c.res_bias = rwt.data_ptr(), self.packed(res_conv[2]).data_ptr()
        self.ops.append(op)
        return out

    def token_linear(self, x: View, w: WSpec, out: View, **kw):
        """Linear layer over a dense token stream (rows, L, C): on tensor-core programs the tokens are flattened to rows*L
        length-1 sequences (any L, e.g. DiT1d's 100), per-trajectory vectors follow through ``sample_row_div = L`` and a
        residual stream is flattened alongside; if the flattened operator is not tensor-core material it is emitted as is."""
        L = x.L
        dense = (x.lstride == x.C and x.bstride == L * x.C and out.lstride == out.C and out.bstride == L * out.C)
        res = kw.get("res")
        res_period = 0                       # > 0: the residual repeats every L rows (a (L, C) table broadcast over the batch)
        if res is not None:
            if res.bstride == 0 and res.lstride == res.C:
                res_period = L
            else:
                dense = dense and res.lstride == res.C and res.bstride == L * res.C
        if self.tc and dense and x.t.dtype == self.act_dtype and w.nk is not None and kw.get("res_conv") is None:
            flat = lambda v: View(v.t, 1, v.C, v.offset, v.C, v.C, v.tf32)
            probe = cabi.Op()
            c = probe.u.conv
            c.batch, c.L_in, c.L_out, c.C_in, c.C_out, c.taps, c.stride, c.pad, c.phases = self.rows * L, 1, 1, x.C, out.C, 1, 1, 0, 1
            c.in_, c.in_bstride, c.in_lstride, c.in_dtype = x.ptr, x.C, x.C, x.dtype
            c.out, c.out_bstride, c.out_lstride, c.out_dtype = out.ptr, out.C, out.C, out.dtype
            c.act = kw.get("act", cabi.ACT_NONE)
            c.sample_row_div = L
            c.in_batch_mod = int(kw.get("in_batch_mod", 0)) * L
            c.math = self.math
            if res is not None:
                c.res, c.res_bstride, c.res_lstride, c.res_dtype = res.ptr, res.C, res.C, res.dtype
                c.res_batch_mod = res_period
            if cabi.load().cds_conv_tc_supported(C.byref(c)):
                kw2 = dict(kw)
                kw2["in_batch_mod"] = c.in_batch_mod
                if res is not None:
                    kw2["res"] = flat(res)
                    kw2["res_batch_mod"] = res_period
                return self.conv(flat(x), w, flat(out), rows=self.rows * L, sample_row_div=L, **kw2)
        return self.conv(x, w, out, **kw)

    def cast_pad(self, x: View, width: int, rows: Optional[int] = None) -> View:
        """fp32 dense (rows, L, C) -> dense (rows, L, width) of the program's activation dtype with zero channels appended
        (rows: of the VIEW, which may be a sub-batch of its tensor)."""
        assert x.t.dtype == torch.float32 and x.lstride == x.C and x.bstride == x.L * x.C
        rows = x.t.shape[0] if rows is None else int(rows)
        out = View(self.buf(rows, x.L, width, dtype=self.act_dtype), x.L, width, tf32=self.math == cabi.MATH_TF32_TC)
        op = cabi.Op()
        op.kind = cabi.OP_CAST
        k = op.u.cast
        k.batch, k.L, k.C_in, k.C_out, k.in_, k.out, k.out_dtype = rows, x.L, x.C, int(width), x.ptr, out.ptr, out.dtype
        self.ops.append(op)
        return out

    def lnmod(self, x: View, out: View, mod: torch.Tensor, shift_col: int, scale_col: int, eps: float):
        op = cabi.Op()
        op.kind = cabi.OP_LNMOD
        m = op.u.lnmod
        m.batch, m.L, m.C, m.eps = self.rows, x.L, x.C, eps
        m.in_, m.out, m.out_dtype = x.ptr, out.ptr, out.dtype
        m.shift, m.scale = mod.data_ptr() + 4 * int(shift_col), mod.data_ptr() + 4 * int(scale_col)
        m.mod_bstride = mod.stride(0)
        self.ops.append(op)

    def attn(self, qkv: View, out: View, heads: int):
        op = cabi.Op()
        op.kind = cabi.OP_ATTN
        a = op.u.attn
        a.batch, a.L, a.C, a.heads = self.rows, out.L, out.C, int(heads)
        a.qkv, a.out, a.out_dtype, a.qkv_dtype = qkv.ptr, out.ptr, out.dtype, qkv.dtype
        self.ops.append(op)


# =============================================================================== UNets
def _is_groupnorm(m) -> bool:
    """GroupNorm1d recognised structurally (class name + the attributes the kernels need), so that modules built from the
    reference's own classes (cleandiffuser/utils/building_blocks.py:60-76) lower exactly like this package's."""
    return (type(m).__name__ in ("GroupNorm1d", "GroupNorm") and hasattr(m, "num_groups") and hasattr(m, "eps")
            and isinstance(getattr(m, "weight", None), torch.Tensor) and isinstance(getattr(m, "bias", None), torch.Tensor))


def _lower_conv_block(p: Program, seq: nn.Sequential, x: View, out: View, k: int, **kw):
    """Conv1d -> GroupNorm1d -> Mish as ONE operator."""
    conv, gn = seq[0], seq[1]
    if not _is_groupnorm(gn):
        raise Unsupported("norm_type other than groupnorm")
    return p.conv(x, w_conv(conv, x.C), out, taps=k, pad=k // 2, bias=_const_vec(p.packed(lambda: conv.bias)),
                  gn=gn, act=cabi.ACT_MISH, **kw)


def _lower_resblock(p: Program, blk, x: View, out: View, k: int, cond: dict):
    """(Chi)ResidualBlock = 2 operators: conv1+GN+Mish+FiLM, then conv2+GN+Mish+shortcut."""
    h = p.act(out.L, out.C)
    _lower_conv_block(p, blk.conv1, x, h, k, **cond)
    if isinstance(blk.residual_conv, nn.Conv1d):
        rc = blk.residual_conv
        shortcut = dict(res_conv=(x, w_rows(lambda: _pad_cin(rc.weight, x.C)[:, :, 0]), lambda: rc.bias))
    else:
        shortcut = dict(res=x)
    return _lower_conv_block(p, blk.conv2, h, out, k, **shortcut)


def _unet_body(p: Program, net, x: View, horizon: int, k: int, film_of: Callable, stage_len: int, final_k: int):
    """Shared down / mid / up walk of both UNets.  ``film_of(block)`` gives the conditioning kwargs.
    Skip connections are never copied: the producer writes straight into one half of the (b, L, 2C)
    buffer the up-stage reads as its concatenated input."""
    n = len(net.downs)
    chans = [net.downs[s][0].conv1[0].out_channels for s in range(n)]
    lens = [horizon >> s for s in range(n)]
    if lens[-1] < 1:
        raise Unsupported("horizon too short for the number of stages")
    # cat buffer of up-stage u holds [x | skip h[n-1-u]] at resolution n-1-u
    cats = [View(p.buf(p.rows, lens[n - 1 - u], 2 * chans[n - 1 - u], dtype=p.act_dtype), lens[n - 1 - u],
                 2 * chans[n - 1 - u], tf32=p.math == cabi.MATH_TF32_TC) for u in range(n - 1)]

    def skip_slot(s):     # where h[s] lives
        if s == 0 or n == 1:
            return p.act(lens[s], chans[s])
        u = n - 1 - s
        return cats[u].channels(chans[s], chans[s])

    for s in range(n):
        stage = net.downs[s]
        mid = _lower_resblock(p, stage[0], x, p.act(lens[s], chans[s]), k, film_of(stage[0]))
        h = _lower_resblock(p, stage[1], mid, skip_slot(s), k, film_of(stage[1]))
        down = stage[stage_len - 1]
        if isinstance(down, nn.Identity):
            x = h
        else:
            x = p.conv(h, w_conv(down.conv), p.act(lens[s + 1], chans[s]), taps=3, stride=2, pad=1,
                       bias=_const_vec(p.packed(lambda d=down: d.conv.bias)))

    mids = [net.mid_block1, net.mid_block2] if hasattr(net, "mid_block1") else list(net.mids)
    x = _lower_resblock(p, mids[0], x, p.act(lens[-1], chans[-1]), k, film_of(mids[0]))
    dst = cats[0].channels(0, chans[-1]) if n > 1 else p.act(lens[-1], chans[-1])
    x = _lower_resblock(p, mids[1], x, dst, k, film_of(mids[1]))

    for u in range(n - 1):
        stage = net.ups[u]
        s = n - 1 - u                          # resolution index of this up stage
        c_out = stage[0].conv1[0].out_channels
        y = _lower_resblock(p, stage[0], cats[u], p.act(lens[s], c_out), k, film_of(stage[0]))
        y = _lower_resblock(p, stage[1], y, p.act(lens[s], c_out), k, film_of(stage[1]))
        up = stage[stage_len - 1]
        if isinstance(up, nn.Identity):
            raise Unsupported("up stage without Upsample1d")
        dst = cats[u + 1].channels(0, c_out) if u + 1 < n - 1 else p.act(lens[s - 1], c_out)
        x = p.conv(y, w_convT(up.conv), dst, taps=3, stride=1, pad=1, phases=2, L_out=lens[s],
                   bias=_const_vec(p.packed(lambda m=up: m.conv.bias)))

    fc = net.final_conv
    y = _lower_conv_block(p, fc, x, p.act(horizon, fc[0].out_channels), final_k)
    last = fc[3]
    return p.conv(y, w_rows(lambda: last.weight[:, :, 0]), p.act(horizon, last.out_channels, dtype=torch.float32),
                  bias=_const_vec(p.packed(lambda: last.bias)))


def _all_resblocks(net):
    blocks = []
    for stage in net.downs:
        blocks += [stage[0], stage[1]]
    blocks += [net.mid_block1, net.mid_block2] if hasattr(net, "mid_block1") else list(net.mids)
    for stage in net.ups:
        blocks += [stage[0], stage[1]]
    return blocks


def lower_janner(p: Program, net: nn.Module, x: View, horizon: int, has_cond: bool, in_batch_mod: int) -> View:
    if any(not isinstance(s[2], nn.Identity) for s in list(net.downs) + list(net.ups)) or \
            not isinstance(net.mid_attn, nn.Identity):
        raise Unsupported("JannerUNet1d(attention=True)")
    if horizon & (horizon - 1):
        raise Unsupported("horizon must be 2^n")
    blocks = _all_resblocks(net)
    widths = [b.emb_mlp[1].out_features for b in blocks]
    offs = [sum(widths[:i]) for i in range(len(widths))]
    total = sum(widths)
    k = blocks[0].conv1[0].kernel_size[0]

    if not has_cond:
        # time conditioning is batch-constant: one row per iteration, evaluated by the model's own modules
        table = p.buf(p.n_iters, total)

        def fill(ctx):
            e = net.map_noise(ctx.t_all)
            e = net.map_emb(e + torch.zeros_like(e))
            table.copy_(torch.cat([b.emb_mlp(e) for b in blocks], dim=1))
        fill.time_only = True      # depends on (weights, timesteps) only: the runtime skips it when neither changed
        p.per_call.append(fill)
        film = {id(b): dict(shift=_vec(step=table, col=o)) for b, o in zip(blocks, offs)}
    else:
        # emb = map_emb(map_noise(t) + cond_b): per trajectory AND per iteration -> tiny GEMMs in the loop.
        md, e_dim = net.map_emb[2].out_features, net.map_emb[0].in_features
        t_rows = p.buf(p.n_iters, net.map_emb[0].out_features)       # W0 map_noise(t) + b0, per iteration
        c_rows = p.buf(p.rows, net.map_emb[0].out_features)          # W0 cond_b, per trajectory
        dummy_in, dummy_w = p.buf(p.rows, 1), p.buf(1, net.map_emb[0].out_features)

        def fill(ctx):
            t_rows.copy_(net.map_emb[0](net.map_noise(ctx.t_all)))
            c_rows.copy_(F.linear(ctx.cond_rows, net.map_emb[0].weight))
        p.per_call.append(fill)
        h1 = View(p.buf(p.rows, 1, net.map_emb[0].out_features), 1, net.map_emb[0].out_features)
        p.conv(View(dummy_in, 1, 1), WSpec(lambda: dummy_w), h1, bias=_vec(step=t_rows, sample=c_rows), act=cabi.ACT_MISH)
        emb_m = View(p.buf(p.rows, 1, md), 1, md)                      # Mish(map_emb(.)): every consumer starts with Mish
        p.conv(h1, w_linear(net.map_emb[2].weight), emb_m, bias=_const_vec(p.packed(lambda: net.map_emb[2].bias)),
               act=cabi.ACT_MISH)
        tb = View(p.buf(p.rows, 1, total), 1, total)
        p.conv(emb_m, w_rows(lambda: torch.cat([b.emb_mlp[1].weight for b in blocks], 0)), tb,
               bias=_const_vec(p.packed(lambda: torch.cat([b.emb_mlp[1].bias for b in blocks], 0))))
        film = {id(b): dict(shift=_vec(sample=tb.t.view(p.rows, total), col=o)) for b, o in zip(blocks, offs)}

    x = View(x.t, x.L, x.C, x.offset, x.bstride, x.lstride, x.tf32)
    # first conv reads x_t; under two-branch CFG both halves of the doubled batch read the same rows
    first = blocks[0]
    pred = _unet_body_with_mod(p, net, x, horizon, k, lambda b: film[id(b)], 4, 5, in_batch_mod, first)
    return pred


def _unet_body_with_mod(p, net, x, horizon, k, film_of, stage_len, final_k, in_batch_mod, first_block):
    """``_unet_body`` with ``in_batch_mod`` applied to the operators that read x_t directly (the first block's
    conv1 and its shortcut)."""
    n_before = len(p.ops)
    pred = _unet_body(p, net, x, horizon, k, film_of, stage_len, final_k)
    if in_batch_mod:
        x_ptr = x.ptr
        for op in p.ops[n_before:]:
            c = op.u.conv
            if op.kind == cabi.OP_CONV and c.in_ == x_ptr:
                c.in_batch_mod = in_batch_mod
            if op.kind == cabi.OP_CONV and (c.res_in == x_ptr or c.res == x_ptr):
                c.res_batch_mod = in_batch_mod
    return pred


def lower_chi(p: Program, net: nn.Module, x: View, horizon: int, has_cond: bool, in_batch_mod: int) -> View:
    if not net.obs_as_global_cond:
        raise Unsupported("ChiUNet1d(obs_as_global_cond=False)")
    if not has_cond:
        raise Unsupported("ChiUNet1d needs a condition")       # the reference raises too (chiunet.py:149)
    if horizon & (horizon - 1):
        raise Unsupported("horizon must be 2^n")
    blocks = _all_resblocks(net)
    widths = [b.cond_encoder[1].out_features for b in blocks]
    offs = [sum(widths[:i]) for i in range(len(widths))]
    total = sum(widths)
    e_dim = net.emb_dim
    k = blocks[0].conv1[0].kernel_size[0]
    step_tab, samp_tab = p.buf(p.n_iters, total), p.buf(p.rows, total)

    def fill(ctx):
        # cond_encoder(cat[time, obs]) = W_t Mish(time) + b  (per iteration)  +  W_o Mish(obs)  (per trajectory)
        te = F.mish(net.map_emb(net.map_noise(ctx.t_all)))
        oe = F.mish(net.global_cond_encoder(torch.flatten(ctx.cond_rows, 1)))
        step_tab.copy_(torch.cat([F.linear(te, b.cond_encoder[1].weight[:, :e_dim], b.cond_encoder[1].bias)
                                  for b in blocks], 1))
        samp_tab.copy_(torch.cat([F.linear(oe, b.cond_encoder[1].weight[:, e_dim:]) for b in blocks], 1))
    p.per_call.append(fill)

    def film_of(b):
        o = offs[[id(z) for z in blocks].index(id(b))]
        if b.cond_predict_scale:
            return dict(scale=_vec(step_tab, samp_tab, o), shift=_vec(step_tab, samp_tab, o + b.out_dim))
        return dict(shift=_vec(step_tab, samp_tab, o))
    return _unet_body_with_mod(p, net, x, horizon, k, film_of, 3, k, in_batch_mod, blocks[0])


# =============================================================================== DQLMlp
def lower_dql(p: Program, net: nn.Module, x: View, has_cond: bool, in_batch_mod: int) -> View:
    act_dim = x.C
    lin1 = net.mid_layer[0]
    e_dim = net.time_mlp[2].out_features
    hidden = lin1.out_features
    step_tab, samp_tab = p.buf(p.n_iters, hidden), p.buf(p.rows, hidden)

    def fill(ctx):
        # Linear over cat[x, time_mlp(map_noise(t)), obs] = W_x x + (W_t temb + b) + W_o obs
        temb = net.time_mlp(net.map_noise(ctx.t_all))
        step_tab.copy_(F.linear(temb, lin1.weight[:, act_dim:act_dim + e_dim], lin1.bias))
        if has_cond:
            samp_tab.copy_(F.linear(ctx.cond_rows, lin1.weight[:, act_dim + e_dim:]))
        else:
            samp_tab.zero_()
    p.per_call.append(fill)
    h = View(p.buf(p.rows, 1, hidden), 1, hidden)
    p.conv(x, w_linear(lin1.weight, slice(0, act_dim)), h, bias=_vec(step_tab, samp_tab), act=cabi.ACT_MISH,
           in_batch_mod=in_batch_mod)
    for idx in (2, 4):
        lin = net.mid_layer[idx]
        h2 = View(p.buf(p.rows, 1, lin.out_features), 1, lin.out_features)
        p.conv(h, w_linear(lin.weight), h2, bias=_const_vec(p.packed(lambda m=lin: m.bias)), act=cabi.ACT_MISH)
        h = h2
    out = View(p.buf(p.rows, 1, act_dim), 1, act_dim)
    fl = net.final_layer
    return p.conv(h, w_linear(fl.weight), out, bias=_const_vec(p.packed(lambda: fl.bias)))


# =============================================================================== IDQLMlp
def lower_idql(p: Program, net: nn.Module, x: View, has_cond: bool, in_batch_mod: int) -> View:
    """idqlmlp.py:21-65.  affine_in over cat[x, time_mlp(t), obs] splits like DQLMlp's first Linear (per-iteration row +
    per-trajectory row); each residual block is LayerNorm (affine folded into the LN+modulate operator: scale = gamma - 1,
    shift = beta, broadcast over the batch) -> Linear + Mish -> Linear + identity residual.  Dropout is inactive in eval mode."""
    if net.training and any(blk.net[0].p > 0 for blk in net.ln_resnet):
        raise Unsupported("IDQLMlp in train mode (dropout)")
    act_dim = x.C
    lin_in = net.affine_in
    e_dim = net.time_mlp[2].out_features
    hidden = lin_in.out_features
    step_tab, samp_tab = p.buf(p.n_iters, hidden), p.buf(p.rows, hidden)

    def fill(ctx):
        temb = net.time_mlp(net.map_noise(ctx.t_all))
        step_tab.copy_(F.linear(temb, lin_in.weight[:, act_dim:act_dim + e_dim], lin_in.bias))
        if has_cond:
            samp_tab.copy_(F.linear(ctx.cond_rows, lin_in.weight[:, act_dim + e_dim:]))
        else:
            samp_tab.zero_()
    p.per_call.append(fill)
    f32 = torch.float32
    h = View(p.buf(p.rows, 1, hidden), 1, hidden)                       # the residual stream stays exact fp32
    p.conv(x, w_linear(lin_in.weight, slice(0, act_dim)), h, bias=_vec(step_tab, samp_tab), in_batch_mod=in_batch_mod)
    for blk in net.ln_resnet:
        ln, fc1, fc2 = blk.net[1], blk.net[2], blk.net[4]
        if not ln.elementwise_affine:
            raise Unsupported("LayerNorm without affine")
        mod = p.buf(1, 2 * hidden)                                    # [gamma - 1 | beta]: LN(x) * (1 + scale) + shift
        p.packers.append(lambda m=mod, l=ln: m.copy_(torch.cat([l.weight - 1., l.bias])[None]))
        p.packers[-1]()
        y = p.act(1, hidden)
        op = cabi.Op()
        op.kind = cabi.OP_LNMOD
        m = op.u.lnmod
        m.batch, m.L, m.C, m.eps = p.rows, 1, hidden, ln.eps
        m.in_, m.out, m.out_dtype = h.ptr, y.ptr, y.dtype
        m.scale, m.shift, m.mod_bstride = mod.data_ptr(), mod.data_ptr() + 4 * hidden, 0
        p.ops.append(op)
        mid = p.act(1, fc1.out_features)
        p.conv(y, w_linear(fc1.weight), mid, bias=_const_vec(p.packed(lambda f=fc1: f.bias)), act=cabi.ACT_MISH)
        h2 = View(p.buf(p.rows, 1, hidden), 1, hidden)
        p.conv(mid, w_linear(fc2.weight), h2, bias=_const_vec(p.packed(lambda f=fc2: f.bias)), res=h)
        h = h2
    out = View(p.buf(p.rows, 1, act_dim), 1, act_dim)
    fo = net.affine_out
    return p.conv(h, w_linear(fo.weight), out, bias=_const_vec(p.packed(lambda: fo.bias)))


# =============================================================================== SfBCUNet
def lower_sfbc(p: Program, net: nn.Module, x: View, has_cond: bool, in_batch_mod: int) -> View:
    """sfbc_unet.py:9-82.  Every residual block is two fused operators:
    ``h = silu(W1 x + b1) + linearc(c)`` (the conditioning enters as the post-activation shift: one row per iteration for the
    time part, one per trajectory for the condition part, both evaluated once per call for all blocks) and
    ``out = silu(W2 h + b2) + skip(x)`` (skip = the operator's shortcut GEMM, or an identity residual).
    ``torch.cat([x, kept.pop()])`` costs nothing: the down block whose output is kept writes straight into the right-hand
    channels of the up block's input buffer, the previous block into the left-hand ones."""
    downs, ups = list(net.down_blocks), list(net.up_blocks)
    blocks = downs + [net.mid_block] + ups
    widths = [b.linearc.out_features for b in blocks]
    offs = [sum(widths[:i]) for i in range(len(widths))]
    total = sum(widths)
    step_tab, samp_tab = p.buf(p.n_iters, total), p.buf(p.rows, total)

    def fill(ctx):
        wc = torch.cat([b.linearc.weight for b in blocks], 0)
        bc = torch.cat([b.linearc.bias for b in blocks], 0)
        step_tab.copy_(F.linear(net.t_layer(net.map_noise(ctx.t_all)), wc, bc))
        if has_cond:
            samp_tab.copy_(F.linear(ctx.cond_rows, wc))
        else:
            samp_tab.zero_()
    p.per_call.append(fill)

    n = len(downs)
    # input buffers of the up blocks: [ x from below | kept down activation ]
    cat_in = []
    for j, b in enumerate(ups):
        left = b.linear1[0].in_features - downs[n - 1 - j].linear1[0].out_features
        cat_in.append((p.act(1, b.linear1[0].in_features), left))

    def out_view(width, dest):
        """Where a block's output goes: a fresh activation, or a channel range of an up block's input buffer."""
        if dest is None:
            return p.act(1, width)
        buf, start = dest
        return buf.channels(start, width)

    def block(b, xin, o, dest, first=False):
        width = b.linear1[0].out_features
        h = p.act(1, width)
        ibm = in_batch_mod if first else 0
        p.conv(xin, w_linear(b.linear1[0].weight), h, bias=_const_vec(p.packed(lambda: b.linear1[0].bias)), act=cabi.ACT_SILU,
               shift=_vec(step=step_tab, sample=samp_tab, col=o), in_batch_mod=ibm)
        out = out_view(width, dest)
        kw = dict(bias=_const_vec(p.packed(lambda: b.linear2[0].bias)), act=cabi.ACT_SILU, res_batch_mod=ibm)
        if isinstance(b.skip, nn.Linear):
            kw["res_conv"] = (xin, w_linear(b.skip.weight), lambda: b.skip.bias)
        else:
            kw["res"] = xin
        return p.conv(h, w_linear(b.linear2[0].weight), out, **kw)

    h = x
    for i, b in enumerate(downs):
        j = n - 1 - i                              # the up block that pops this activation (none for the first down block)
        dest = (cat_in[j][0], cat_in[j][1]) if 0 <= j < len(ups) else None
        h = block(b, h, offs[i], dest, first=(i == 0))
    h = block(net.mid_block, h, offs[n], (cat_in[0][0], 0) if ups else None)
    for j, b in enumerate(ups):
        dest = (cat_in[j + 1][0], 0) if j + 1 < len(ups) else None
        h = block(b, cat_in[j][0], offs[n + 1 + j], dest)
    out = View(p.buf(p.rows, 1, x.C), 1, x.C)
    return p.conv(h, w_linear(net.out_layer.weight), out, bias=_const_vec(p.packed(lambda: net.out_layer.bias)))


# =============================================================================== PearceMlp
def lower_pearce(p: Program, net: nn.Module, x: View, has_cond: bool, in_batch_mod: int) -> View:
    """pearcemlp.py:35-79.  Every Linear over a concatenation splits by column blocks, like DQLMlp's first layer:
      fcs[0]  W [x_e | t_e | cond]      = W_xe x_e  + (W_te map_noise(t) + b: one row per iteration) + (W_c cond: one row per trajectory)
      fcs[k]  W [nn/1.414 | x | t]      = (W_h / c) v + (x W_x^T: one row per trajectory AND iteration, a small GEMM ahead of the
                                          blocks, all three layers at once) + (w_t t + b: one row per iteration)
    The residual chain ``nn_{k+1} = FC_k(nn_k / 1.414, ..) + nn_k / 1.414`` is carried as ``v_k = 1.414^(k-1) nn_k`` so that every
    residual has coefficient one: ``v_{k+1} = 1.414^k FC_k + v_k`` (a constant output scale), the 1 / 1.414^k go into the packed
    weights of the Linear that reads ``v_k``.  FCBlock = Linear + GroupNorm1d + exact GELU = one operator."""
    act_dim, emb, To = x.C, net.emb_dim, net.To
    fc0, fc1, fc2, head = net.fcs[0], net.fcs[1], net.fcs[2], net.fcs[3]
    for fc in (fc0, fc1, fc2):
        if not _is_groupnorm(fc.model[1]) or type(fc.model[2]).__name__ != "GELU" or getattr(fc.model[2], "approximate", "none") != "none":
            raise Unsupported("PearceMlp FCBlock other than Linear -> GroupNorm1d -> GELU")
    hidden = fc0.model[0].out_features
    lins = [fc1.model[0], fc2.model[0], head]
    s = 1.414
    inv = [1.0 / s, 1.0 / (s * s), 1.0 / (s * s)]       # what the Linear reading v_1, v_2, v_3 folds into its weights
    widths = [l.out_features for l in lins]
    offs = [sum(widths[:i]) for i in range(3)]
    xw = p.buf(p.rows, sum(widths))                                     # x W_x^T of the three layers
    step0, samp0 = p.buf(p.n_iters, hidden), p.buf(p.rows, hidden)
    steps = [p.buf(p.n_iters, w) for w in widths]

    def fill(ctx):
        w0 = fc0.model[0]
        step0.copy_(F.linear(net.map_noise(ctx.t_all), w0.weight[:, emb:2 * emb], w0.bias))
        if has_cond:
            samp0.copy_(F.linear(torch.flatten(ctx.cond_rows, 1), w0.weight[:, 2 * emb:]))
        else:
            samp0.zero_()
        tr = ctx.t_all.to(torch.float32)[:, None]
        for lin, tab in zip(lins, steps):
            tab.copy_(tr * lin.weight[:, hidden + act_dim][None, :] + lin.bias[None, :])
    p.per_call.append(fill)

    p.conv(x, w_rows(lambda: torch.cat([l.weight[:, hidden:hidden + act_dim] for l in lins], 0)), View(xw, 1, xw.shape[1]),
           in_batch_mod=in_batch_mod)
    ae = net.act_emb
    e1 = p.act(1, ae[0].out_features)
    p.conv(x, w_linear(ae[0].weight), e1, bias=_const_vec(p.packed(lambda: ae[0].bias)), act=cabi.ACT_LEAKY_RELU, in_batch_mod=in_batch_mod)
    x_e = p.act(1, emb)
    p.conv(e1, w_linear(ae[2].weight), x_e, bias=_const_vec(p.packed(lambda: ae[2].bias)))
    f32 = torch.float32
    v = View(p.buf(p.rows, 1, hidden), 1, hidden)                        # the residual chain stays exact fp32
    p.conv(x_e, w_linear(fc0.model[0].weight, slice(0, emb)), v, bias=_vec(step0, samp0), gn=fc0.model[1], act=cabi.ACT_GELU_ERF)
    for k, fc in enumerate((fc1, fc2)):
        lin = fc.model[0]
        v2 = View(p.buf(p.rows, 1, hidden), 1, hidden)
        scale = p.buf(1, hidden)
        scale.fill_(s ** (k + 1))
        p.conv(v, w_rows(lambda l=lin, c=inv[k]: l.weight[:, :hidden] * c), v2, bias=cabi_vec2(steps[k], xw, offs[k]), gn=fc.model[1], act=cabi.ACT_GELU_ERF, scale=_const_vec(scale[0]), res=v)
        v = v2
    out = View(p.buf(p.rows, 1, act_dim), 1, act_dim)
    return p.conv(v, w_rows(lambda: head.weight[:, :hidden] * inv[2]), out, bias=cabi_vec2(steps[2], xw, offs[2]))


def cabi_vec2(step: torch.Tensor, sample: torch.Tensor, sample_col: int) -> cabi.Vec:
    """step: [n_iters, W] table from column 0; sample: [rows, Wtot] table from column ``sample_col``."""
    v = cabi.Vec()
    v.step, v.step_stride = step.data_ptr(), step.stride(0)
    v.sample, v.sample_stride = sample.data_ptr() + 4 * int(sample_col), sample.stride(0)
    return v


# =============================================================================== DiT1d
def lower_dit(p: Program, net: nn.Module, x: View, horizon: int, has_cond: bool, in_batch_mod: int) -> View:
    d = net.d_model
    depth = len(net.blocks)
    heads = net.blocks[0].attn.num_heads
    if net.blocks[0].attn.dropout != 0.0 and net.training:
        raise Unsupported("attention dropout in train mode")
    R, L = p.rows, horizon

    # ---- conditioning: emb = Mish(W2 Mish(W0 (map_noise(t) + cond_b) + b0) + b2);  every consumer applies SiLU first
    w0 = net.map_emb[0]
    t_rows, c_rows = p.buf(p.n_iters, d), p.buf(R, d)
    pos = p.buf(L, d)
    dummy_in, dummy_w = p.buf(R, 1), p.buf(1, d)

    def fill(ctx):
        t_rows.copy_(w0(net.map_noise(ctx.t_all)))
        if has_cond:
            c_rows.copy_(F.linear(ctx.cond_rows, w0.weight))
        else:
            c_rows.zero_()
        pos.copy_(net.pos_emb(torch.arange(L, device=pos.device)))      # int64 positions: degenerate table, by design
    p.per_call.append(fill)
    h1 = View(p.buf(R, 1, d), 1, d)
    p.conv(View(dummy_in, 1, 1), WSpec(lambda: dummy_w), h1, bias=_vec(step=t_rows, sample=c_rows), act=cabi.ACT_MISH)
    emb_s = View(p.buf(R, 1, d), 1, d)
    p.conv(h1, w_linear(net.map_emb[2].weight), emb_s, bias=_const_vec(p.packed(lambda: net.map_emb[2].bias)),
           act=cabi.ACT_MISH_SILU)
    mods = [blk.adaLN_modulation[1] for blk in net.blocks] + [net.final_layer.adaLN_modulation[1]]
    total = sum(m.out_features for m in mods)
    mod = View(p.buf(R, 1, total), 1, total)
    p.conv(emb_s, w_rows(lambda: torch.cat([m.weight for m in mods], 0)), mod,
           bias=_const_vec(p.packed(lambda: torch.cat([m.bias for m in mods], 0))))
    mod2d = mod.t.view(R, total)

    # ---- tokens
    # residual stream X and the attention operands QKV stay fp32; what feeds a Linear layer (the modulated tokens Y, the
    # attention output, the MLP hidden) has the program's activation dtype: bf16 on tensor-core programs, where those Linear
    # layers (97 % of DiT1d's FLOPs outside attention) run on tcgen05 over the flattened token stream
    f32 = torch.float32
    X = p.act(L, d, f32)
    # head_dim 32 (every pipeline) and L <= 128: attention runs on tensor cores (mma.sync, bf16 q/k/v); otherwise fp32 CUDA cores
    QKV = p.act(L, 3 * d) if (d // heads == 32 and L <= 128) else p.act(L, 3 * d, f32)
    Y, ATT, HID = p.act(L, d), p.act(L, d), p.act(L, 4 * d)
    if p.tc and x.t.dtype == torch.float32 and x.lstride == x.C and x.bstride == L * x.C:
        # x_t enters as a channel-padded copy in the activation dtype (kept fresh by the solver update, like the UNets' hand-over)
        kin = p.k_align * ((x.C + p.k_align - 1) // p.k_align)
        xb = p.cast_pad(x, kin, rows=in_batch_mod or p.rows)
        w_in = w_rows(lambda: F.pad(net.x_proj.weight, (0, kin - net.x_proj.weight.shape[1])))
        p.token_linear(xb, w_in, X, bias=_const_vec(p.packed(lambda: net.x_proj.bias)),
                       res=View(pos, L, d, bstride=0), in_batch_mod=in_batch_mod)
    else:
        p.conv(x, w_linear(net.x_proj.weight), X, bias=_const_vec(p.packed(lambda: net.x_proj.bias)),
               res=View(pos, L, d, bstride=0), in_batch_mod=in_batch_mod)
    for i, blk in enumerate(net.blocks):
        o = 6 * d * i       # chunk order: shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
        p.lnmod(X, Y, mod2d, o, o + d, blk.norm1.eps)
        at = blk.attn
        p.token_linear(Y, w_linear(at.in_proj_weight), QKV, bias=_const_vec(p.packed(lambda a=at: a.in_proj_bias)))
        p.attn(QKV, ATT, heads)
        # reference quirk (dit.py:33-34): the residual wraps the MODULATED tokens Y, not the block input
        p.token_linear(ATT, w_linear(at.out_proj.weight), X, bias=_const_vec(p.packed(lambda a=at: a.out_proj.bias)),
                       scale=_vec(sample=mod2d, col=o + 2 * d), res=Y)
        p.lnmod(X, Y, mod2d, o + 3 * d, o + 4 * d, blk.norm2.eps)
        p.token_linear(Y, w_linear(blk.mlp[0].weight), HID, bias=_const_vec(p.packed(lambda b=blk: b.mlp[0].bias)),
                       act=cabi.ACT_GELU_TANH)
        p.token_linear(HID, w_linear(blk.mlp[3].weight), X, bias=_const_vec(p.packed(lambda b=blk: b.mlp[3].bias)),
                       scale=_vec(sample=mod2d, col=o + 5 * d), res=X)
    o = 6 * d * depth
    fl = net.final_layer
    p.lnmod(X, Y, mod2d, o, o + d, fl.norm_final.eps)
    out = p.act(L, net.in_dim, f32)
    return p.token_linear(Y, w_linear(fl.linear.weight), out, bias=_const_vec(p.packed(lambda: fl.linear.bias)))


def lower_denoiser(p: Program, net: nn.Module, x: View, x_shape, has_cond: bool, in_batch_mod: int) -> View:
    """Dispatch on the backbone type (reference instances are recognised structurally by class name)."""
    name = type(net).__name__
    if name in ("JannerUNet1d", "ChiUNet1d") and len(x_shape) == 2 and p.tc:
        # TMA/UMMA want K in whole 64-byte rows: 32 bf16 / 16 fp32 channels
        x = p.cast_pad(x, p.k_align * ((x.C + p.k_align - 1) // p.k_align), rows=in_batch_mod or p.rows)
    if name == "JannerUNet1d" and len(x_shape) == 2:
        return lower_janner(p, net, x, x_shape[0], has_cond, in_batch_mod)
    if name == "ChiUNet1d" and len(x_shape) == 2:
        return lower_chi(p, net, x, x_shape[0], has_cond, in_batch_mod)
    if name == "DiT1d" and len(x_shape) == 2:
        return lower_dit(p, net, x, x_shape[0], has_cond, in_batch_mod)
    if name == "DQLMlp" and len(x_shape) == 1:
        return lower_dql(p, net, x, has_cond, in_batch_mod)
    if name == "DVInvMlp" and len(x_shape) == 1:
        if not has_cond:
            raise Unsupported("DVInvMlp needs a condition")         # the reference raises too (dvinvmlp.py:44)
        return lower_dql(p, net, x, has_cond, in_batch_mod)         # same graph: cat[x, time_mlp(t), cond] -> 3 x (Linear + Mish) -> Linear
    if name == "SfBCUNet" and len(x_shape) == 1:
        return lower_sfbc(p, net, x, has_cond, in_batch_mod)
    if name == "PearceMlp" and len(x_shape) == 1:
        return lower_pearce(p, net, x, has_cond, in_batch_mod)
    if name == "IDQLMlp" and len(x_shape) == 1:
        return lower_idql(p, net, x, has_cond, in_batch_mod)
    raise Unsupported(f"backbone {name} with x_shape {tuple(x_shape)}")
