"""Backward launch program of the denoiser (the reverse walk of engine.DenoiserEngine.forward).

The reference obtains these gradients from torch.autograd over ddpm.py:477-505; here every
activation gradient is produced by a C-ABI kernel over halo grids:

* convolution input gradients reuse the *forward* convolution kernels with tap-reversed, transposed
  weights over the (zero-halo) output gradient, storing every row; the halo rows of the result are
  then folded onto the border voxels (``tdb_halo_fold`` = adjoint of the replicate halo);
* weight gradients: ``tdb_conv3d_wgrad`` (fp32 accumulation), ``tdb_cl_nc_outer`` for the 1x1x1
  encoders / decoder;
* GroupNorm + FiLM + SiLU: ``tdb_pointwise_bwd_reduce`` / ``_apply``; trilinear: ``tdb_trilinear_bwd``;
  attention: ``tdb_attention_bwd``.

Only O(B*C)-sized bookkeeping (turning the per-channel sums into parameter gradients, the timestep
MLP with its 32..128-wide matrices) is done with torch ops on tiny tensors.
"""

from __future__ import annotations

import torch

from . import _lib
from ._lib import PW_NOHALO, PW_SILU, call, ptr
from .engine import GN_EPS, View


def grad_phase(name: str) -> int:
    """1: the gradient of parameter `name` is complete after BackwardProgram.phase1 (decoder, up path, centre);
    2: after phase2 (down path, encoders, and everything fed by the timestep conditioning - its gradient sums over all
    blocks).  Data-parallel training exchanges the phase-1 gradients while phase 2 is still running."""
    late = ("project_onto_scale_shift" in name or name.startswith("process_c") or name.startswith("encode_")
            or "downsampling_blocks" in name)
    return 2 if late else 1


class BackwardProgram:
    def __init__(self, eng):
        self.eng = eng
        self.m = eng.model
        self.side = None
        self._readers = {}

    # ------------------------------------------------------------------ helpers
    def _gbuf(self, p, v: View, key) -> View:
        """Gradient buffer with the geometry of forward view `v` (cached per plan)."""
        gb = p.setdefault("gbuf", {})
        if key not in gb:
            gb[key] = p["grid"](v.level, v.t.shape[-1])
        base = gb[key]
        return View(base.t, v.c0, v.C, v.level)

    def _tmp(self, p, level, C, key) -> View:
        gb = p.setdefault("gbuf", {})
        k = ("tmp", key, level, C)  # exact channel pitch: half-used 128-byte lines double the DRAM traffic
        if k not in gb:
            gb[k] = p["grid"](level, C)
        return gb[k]

    def _dgrad_weights(self, conv, key, level):
        """Kernel-layout weights of the input-gradient convolution: W'[ci][co][k] = W[co][ci][2-k]."""
        eng = self.eng
        cache = eng._wcache.setdefault("dgrad", {})
        if key not in cache:
            cache[key] = eng.pack_conv(conv.weight.detach(), level, dgrad=True)  # (Cin, Cout, k, k, k): "Cout'" = Cin, "Cin'" = Cout
        return cache[key]

    def _fold(self, p, g: View):
        X, Y, Z = p["sizes"][g.level]
        call("tdb_halo_fold", g.ptr, g.ld, p["B"], X, Y, Z, g.C, self.eng.dt, _lib.stream_ptr())

    def _wgrad(self, p, x: View, d_out: View, conv, ntaps, zero_halo=False):
        """zero_halo: d_out comes from tdb_pointwise_bwd_apply (halo rows are zero) - lets the bf16 path use tensor cores.
        Weight gradients are leaves of the backward walk (nothing downstream reads them before the optimizer), so they
        run on a side stream: the tensor-bound tcgen05 kernel then overlaps the HBM-bound halo fold / GroupNorm-SiLU
        backward passes of the next layer on the main stream.  The buffers it reads are protected by events
        (_wait_readers before they are overwritten); run() joins the side stream at the end."""
        X, Y, Z = p["sizes"][x.level]
        k = 3 if ntaps == 27 else 1

        def launch():
            dw = self._dw_zeros(ntaps * x.C * d_out.C, x.t.device)
            call("tdb_conv3d_wgrad", x.ptr, x.ld, d_out.ptr, d_out.ld, dw.data_ptr(), p["B"], X, Y, Z, x.C, d_out.C, ntaps, self.eng.dt,
                 _lib.WGRAD_ZERO_HALO if zero_halo else 0, _lib.stream_ptr())
            out = torch.empty((d_out.C, x.C, k, k, k), dtype=torch.float32, device=x.t.device)
            call("tdb_unpack_wgrad", dw.data_ptr(), out.data_ptr(), d_out.C, x.C, ntaps, _lib.stream_ptr())  # -> (Cout, Cin, k, k, k)
            return out

        side = self.side
        if side is None:
            return launch()
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            res = launch()
            done = torch.cuda.Event()
            done.record(side)
        if not torch.cuda.is_current_stream_capturing():
            res.record_stream(main)  # (inside a capture every result lives until the join at the end of run())
        self._readers[d_out.t.data_ptr()] = done
        return res

    def _dw_zeros(self, n: int, dev):
        """n zeroed floats for one weight gradient in the kernels' accumulation layout, out of ONE buffer per backward pass
        (zeroed by a single fill on the stream of the first weight gradient instead of a torch.zeros launch in front of each)."""
        pool = getattr(self, "_dw_pool", None)
        if pool is None or self._dw_used + n > pool.numel():
            m = self.m
            convs = [bp.blk.block1.conv for bp in self.eng.blocks.values()] + [bp.blk.block2.conv for bp in self.eng.blocks.values()]
            convs += [bp.blk.conv for bp in self.eng.blocks.values() if bp.has_proj]
            att = m.u_net.center_block[1].fn.fn
            convs += [att.to_qkv, att.to_out]
            total = sum(-(-c.weight.numel() // 64) * 64 for c in convs)
            pool = self._dw_pool = torch.zeros(max(total, n), dtype=torch.float32, device=dev)
            self._dw_used = 0
        out = pool[self._dw_used : self._dw_used + n]
        self._dw_used += -(-n // 64) * 64  # 256-byte aligned slices
        return out

    def _leaf(self, fn):
        """Run a gradient leaf (reads only tensors that nothing overwrites before the end of run()) on the side stream."""
        side = self.side
        if side is None:
            return fn()
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            res = fn()
        if not torch.cuda.is_current_stream_capturing():
            for t in res:
                t.record_stream(main)
        return res

    def _zeros64(self, n: int, dev):
        """n zeroed doubles out of ONE buffer per backward pass (a torch.zeros per norm layer is a fill launch on the critical
        path in front of each of the 23 reduce kernels)."""
        pool = getattr(self, "_zpool", None)
        if pool is None or self._zused + n > pool.numel():
            B = self.p["B"]
            total = B * 4 * (sum(2 * bp.cout for bp in self.eng.blocks.values()) + 2 * self.m.dim * 2 ** self.m.u_net_levels)
            pool = self._zpool = torch.zeros(max(total, n), dtype=torch.float64, device=dev)
            self._zused = 0
        out = pool[self._zused : self._zused + n]
        self._zused += n
        return out

    def _wait_readers(self, v: View):
        """Before `v` is overwritten on the main stream: wait for the side-stream kernel that still reads it."""
        ev = self._readers.pop(v.t.data_ptr(), None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def _colsum(self, p, g: View):
        """Per-channel sum over interior voxels and samples (conv bias gradients)."""
        X, Y, Z = p["sizes"][g.level]
        acc = torch.zeros((p["B"], g.C, 2), dtype=torch.float64, device=g.t.device)
        call("tdb_gn_stats", g.ptr, g.ld, acc.data_ptr(), p["B"], X, Y, Z, g.C, g.C, self.eng.dt, _lib.stream_ptr())
        return acc[:, :, 0].sum(0).float()

    def _pw_bwd(self, p, g_out: View, raw: View, stats, norm, film_view, d_raw: View, flags, G, d_film=None, film_offset=0):
        """Backward of tdb_pointwise (GroupNorm + FiLM + SiLU).  Writes d_raw (zero halo) and, when d_film is given, the
        FiLM scale/shift gradients of this block into d_film[:, film_offset : film_offset + 2C].  Returns fp32 [C]
        tensors (norm weight gradient, norm bias gradient, colsum) where colsum = sum of d_raw over samples and interior
        voxels = the bias gradient of the convolution that produced `raw` (no extra pass over d_raw)."""
        eng = self.eng
        X, Y, Z = p["sizes"][raw.level]
        B, C = p["B"], raw.C
        dev = raw.t.device
        film_ptr = None if film_view is None else film_view.data_ptr()
        red = self._zeros64(B * C * 4, dev).view(B, C, 4)
        call("tdb_pointwise_bwd_reduce", g_out.ptr, g_out.ld, raw.ptr, raw.ld, ptr(stats), ptr(norm.weight), ptr(norm.bias), film_ptr,
             eng.film_rows, red.data_ptr(), B, X, Y, Z, C, G, GN_EPS, flags, eng.dt, _lib.stream_ptr())
        out = torch.empty((4, C), dtype=torch.float32, device=dev)
        dfilm_ptr = None if d_film is None else d_film.data_ptr() + 4 * film_offset
        self._wait_readers(d_raw)
        # group sums + parameter gradients + input gradient in ONE launch (the one-block finalize kernel between reduce and
        # apply was a serial chain of dependent loads on the critical path: 1.3 ms per step for its 23 launches)
        call("tdb_pointwise_bwd_apply_fused", g_out.ptr, g_out.ld, raw.ptr, raw.ld, ptr(stats), ptr(norm.weight), ptr(norm.bias), film_ptr,
             eng.film_rows, red.data_ptr(), d_raw.ptr, d_raw.ld, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
             dfilm_ptr, eng.film_rows, B, X, Y, Z, C, G, GN_EPS, flags, eng.dt, _lib.stream_ptr())
        self._g_colsum = out[3]  # sum of g_out per channel: bias gradient of a 1x1 residual projection on the same output
        return out[1], out[2], out[0]

    def _add_interior(self, p, a: View, b: View):
        """a[interior] += b[interior] (halo rows of `a` untouched)."""
        eng = self.eng
        X, Y, Z = p["sizes"][a.level]
        # the streaming kernels take at most 256 channel vectors of 16 bytes per voxel (2048 bf16 / 1024 fp32 channels): the
        # 2048-channel input gradient of the first up block of a dim = 64 model is added in slices on the fp32 path
        step = 256 * 16 // a.t.element_size()
        for c0 in range(0, a.C, step):
            n = min(step, a.C - c0)
            sa, sb = a.slice(c0, n), b.slice(c0, n)
            call("tdb_pointwise", sa.ptr, sa.ld, None, None, None, None, 0, sb.ptr, sb.ld, sa.ptr, sa.ld, p["B"], X, Y, Z, n, 1, GN_EPS,
                 PW_NOHALO, eng.dt, _lib.stream_ptr())

    # ------------------------------------------------------------------ blocks
    def _resblock_bwd(self, p, name, g_out: View, grads: dict, d_film: torch.Tensor):
        """g_out: folded gradient of the block output.  Returns the folded gradient of the block input."""
        eng = self.eng
        bp = eng.blocks[name]
        blk = bp.blk
        sv = p["saved"][name]
        x, slot = sv["x"], sv["slot"]
        lvl = x.level
        C = bp.cout
        G = eng._groups(C)
        pre = self.prefix[name]
        film = p["film"][:, bp.film_offset : bp.film_offset + 2 * C]
        # two d_raw buffers per (level, channels): block1's pointwise backward writes the other one while block2's weight
        # gradient may still be reading the first on the side stream
        d_raw = self._tmp(p, lvl, C, "d_raw")
        d_raw1 = self._tmp(p, lvl, C, "d_raw1") if self.side is not None else d_raw
        g_act = self._tmp(p, lvl, C, "g_act")

        # block2: pointwise (norm, SiLU, + residual) then conv2
        gw, gb, bias2 = self._pw_bwd(p, g_out, sv["raw2"], p["stats"][slot + 1], blk.block2.norm, None, d_raw, PW_SILU, G)
        g_out_colsum = self._g_colsum
        grads[f"{pre}.block2.norm.weight"] = gw
        grads[f"{pre}.block2.norm.bias"] = gb
        grads[f"{pre}.block2.conv.bias"] = bias2
        # the input gradient (main stream, critical path) is enqueued BEFORE the weight gradient (side stream): both want
        # every SM, so the weight gradient starts as the input gradient's CTAs retire and then overlaps the bandwidth-bound
        # fold / GroupNorm-SiLU backward of the next layer
        eng._conv(p, d_raw, self._dgrad_weights(blk.block2.conv, f"{name}.conv2", lvl), None, g_act, 27, all_rows=True)
        grads[f"{pre}.block2.conv.weight"] = self._wgrad(p, sv["act1"], d_raw, blk.block2.conv, 27, zero_halo=True)
        self._fold(p, g_act)

        # block1: pointwise (norm, FiLM, SiLU) then conv1; the FiLM scale/shift gradients go straight into d_film
        d_raw = d_raw1
        gw, gb, bias1 = self._pw_bwd(p, g_act, sv["raw1"], p["stats"][slot], blk.block1.norm, film, d_raw, PW_SILU, G,
                                     d_film=d_film, film_offset=bp.film_offset)
        grads[f"{pre}.block1.norm.weight"] = gw
        grads[f"{pre}.block1.norm.bias"] = gb
        grads[f"{pre}.block1.conv.bias"] = bias1
        g_x = self._gbuf(p, x, ("g", name))
        # residual projection: its input gradient res_conv^T(g_out) rides on conv1's input-gradient kernel when that is the
        # row-window pair kernel (g_out is a folded gradient - zero halo rows - so the sum may be folded afterwards)
        fuse_res = bp.has_proj and eng.can_add1x1(C, x.C, lvl)
        eng._conv(p, d_raw, self._dgrad_weights(blk.block1.conv, f"{name}.conv1", lvl), None, g_x, 27, all_rows=True,
                  add1x1=(g_out, self._dgrad_weights(blk.conv, f"{name}.proj", lvl)) if fuse_res else None)
        grads[f"{pre}.block1.conv.weight"] = self._wgrad(p, x, d_raw, blk.block1.conv, 27, zero_halo=True)
        self._fold(p, g_x)

        # residual branch
        if bp.has_proj:
            # g_out is a folded gradient: tdb_halo_fold left its halo rows zero
            grads[f"{pre}.conv.weight"] = self._wgrad(p, x, g_out, blk.conv, 1, zero_halo=True)
            grads[f"{pre}.conv.bias"] = g_out_colsum  # from block2's reduction over the same g_out (no extra pass)
            if not fuse_res:
                g_res = self._tmp(p, lvl, x.C, "g_res")
                eng._conv(p, g_out, self._dgrad_weights(blk.conv, f"{name}.proj", lvl), None, g_res, 1, all_rows=True)
                self._add_interior(p, g_x, g_res)
        else:
            self._add_interior(p, g_x, g_out)
        return g_x

    def _attention_bwd(self, p, g_out: View, grads: dict):
        eng, m = self.eng, self.m
        pre_mod = m.u_net.center_block[1].fn
        att = pre_mod.fn
        x = p["center0"]
        X, Y, Z = p["sizes"][x.level]
        B = p["B"]
        s = _lib.stream_ptr
        pre = "u_net.center_block.1.fn"
        G = eng._groups(x.C)
        # to_out (1x1 conv + bias) on the attention output
        grads[f"{pre}.fn.to_out.weight"] = self._wgrad(p, p["attn_o"], g_out, att.to_out, 1)
        grads[f"{pre}.fn.to_out.bias"] = self._colsum(p, g_out)
        g_o = self._gbuf(p, p["attn_o"], ("g", "attn_o"))
        eng._conv(p, g_out, self._dgrad_weights(att.to_out, "attn.out", x.level), None, g_o, 1, all_rows=True)
        # softmax attention
        g_qkv = self._gbuf(p, p["attn_qkv"], ("g", "attn_qkv"))
        call("tdb_attention_bwd", p["attn_qkv"].ptr, p["attn_qkv"].ld, g_o.ptr, g_o.ld, g_qkv.ptr, g_qkv.ld, B, X, Y, Z, att.heads,
             att.dim_head, eng.dt, s())
        # to_qkv (1x1 conv, no bias) on the normalised input
        grads[f"{pre}.fn.to_qkv.weight"] = self._wgrad(p, p["attn_norm"], g_qkv, att.to_qkv, 1)
        g_hn = self._gbuf(p, p["attn_norm"], ("g", "attn_norm"))
        eng._conv(p, g_qkv, self._dgrad_weights(att.to_qkv, "attn.qkv", x.level), None, g_hn, 1, all_rows=True)
        # pre-norm
        g_x = self._gbuf(p, x, ("g", "attn_x"))
        gw, gb, _ = self._pw_bwd(p, g_hn, x, p["stats"][p["attn_slot"]], pre_mod.norm, None, g_x, 0, G)
        grads[f"{pre}.norm.weight"] = gw
        grads[f"{pre}.norm.bias"] = gb
        # residual: out = proj + x
        self._add_interior(p, g_x, g_out)
        return g_x

    # ------------------------------------------------------------------ whole network
    @torch.no_grad()
    def run(self, g_eps: torch.Tensor):
        """Both phases back to back: (parameter gradients by name, gradient of c_local)."""
        self.phase1(g_eps)
        return self.phase2()

    def _join(self):
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)  # every weight gradient issued so far is complete
            self._readers.clear()

    def phase1(self, g_eps: torch.Tensor):
        """Decoder, up path and centre: when it returns, every gradient of `grad_phase(name) == 1` parameters is complete
        (data-parallel training exchanges them while phase 2 runs)."""
        eng, m = self.eng, self.m
        key = eng._last_train_key
        if key is None or key not in eng._plans:
            raise RuntimeError("turbdiff_b200: backward called without a preceding forward(train=True)")
        p = self.p = eng._plans[key]
        x_in, t, c_local = p["last_input"]
        B, F = x_in.shape[:2]
        X, Y, Z = x_in.shape[2:]
        L = m.u_net_levels
        dev = x_in.device
        s = _lib.stream_ptr
        dt = eng.dt
        eng._set_geometry(key[1])
        eng.weights()
        self.side = eng.side_stream(dev) if eng.wgrad_side_stream else None
        self._readers = {}
        self._zpool = None
        self._dw_pool = None
        g_eps = g_eps.to(torch.float32).contiguous()
        grads = self.grads = {}
        d_film = self.d_film = torch.zeros((B, eng.film_rows), dtype=torch.float32, device=dev)
        self.prefix = {"decode0": "decode.0", "center0": "u_net.center_block.0", "center2": "u_net.center_block.2"}
        for i in range(L):
            self.prefix[f"down{i}"] = f"u_net.downsampling_blocks.{i}"
            self.prefix[f"up{i}"] = f"u_net.upsampling_blocks.{i}"

        # decode[1]: 1x1 conv dim -> F read from NCDHW eps gradient
        dec = m.decode[1]
        dec_out = p["dec_out"]
        Fo = m.out_features  # 2F with learned variances

        def dec_grads():
            dw = torch.zeros((m.dim, Fo), dtype=torch.float32, device=dev)
            call("tdb_cl_nc_outer", dec_out.ptr, dec_out.ld, g_eps.data_ptr(), Fo * X * Y * Z, dw.data_ptr(), None, B, X, Y, Z, m.dim, Fo, dt, s())
            return dw.t().reshape(dec.weight.shape).contiguous(), g_eps.sum(dim=(0, 2, 3, 4))

        grads["decode.1.weight"], grads["decode.1.bias"] = self._leaf(dec_grads)
        g_dec = self._gbuf(p, dec_out, ("g", "dec_out"))
        wt = dec.weight.detach().reshape(Fo, m.dim).t().contiguous()  # (dim, Fo)
        zero_b = torch.zeros(m.dim, dtype=torch.float32, device=dev)
        call("tdb_encode_input", g_eps.data_ptr(), None, wt.data_ptr(), zero_b.data_ptr(), None, None, g_dec.ptr, g_dec.ld, B, Fo, 0,
             m.dim, X, Y, Z, 1, dt, s())  # halo rows hold copies, never read: the 1x1 conv only saw interior voxels

        g = self._resblock_bwd(p, "decode0", g_dec, grads, d_film)
        # up path (reverse): block input = cat[l] = [upsampled | skip]
        for i in reversed(range(L)):
            l = L - 1 - i
            g_cat = self._resblock_bwd(p, f"up{i}", g, grads, d_film)  # folded, 2C channels
            C = g_cat.C // 2
            src = p["up_out"][l + 1] if l + 1 < L else p["center2"]
            g = self._gbuf(p, src, ("g", "up_src", l))
            Xo, Yo, Zo = p["sizes"][l]
            Xi, Yi, Zi = p["sizes"][l + 1]
            up_half = g_cat.slice(0, C)
            call("tdb_trilinear_bwd", up_half.ptr, up_half.ld, Xo, Yo, Zo, g.ptr, g.ld, Xi, Yi, Zi, B, C, dt, 0, s())
            p.setdefault("g_skip", {})[l] = g_cat.slice(C, C)
        # centre
        g = self._resblock_bwd(p, "center2", g, grads, d_film)
        g = self._attention_bwd(p, g, grads)
        self.g = self._resblock_bwd(p, "center0", g, grads, d_film)
        self._join()

    def phase2(self):
        """Down path, encoders, timestep MLP / FiLM projections.  Returns (all parameter gradients by name, gradient of c_local)."""
        eng, m, p, grads, d_film, g = self.eng, self.m, self.p, self.grads, self.d_film, self.g
        x_in, t, c_local = p["last_input"]
        B, F = x_in.shape[:2]
        X, Y, Z = x_in.shape[2:]
        L = m.u_net_levels
        dev = x_in.device
        s = _lib.stream_ptr
        dt = eng.dt
        # down path (reverse): skip gradient + gradient through the downsampling
        for l in reversed(range(L)):
            g_skip = p["g_skip"][l]
            Xi, Yi, Zi = p["sizes"][l]
            Xo, Yo, Zo = p["sizes"][l + 1]
            # skip gradient += gradient through the down-sampling, in one pass (no temporary grid, no separate add)
            call("tdb_trilinear_bwd", g.ptr, g.ld, Xo, Yo, Zo, g_skip.ptr, g_skip.ld, Xi, Yi, Zi, B, g_skip.C, dt, _lib.TRIBWD_ACCUMULATE, s())
            g = self._resblock_bwd(p, f"down{l}", g_skip, grads, d_film)
        # encoders (g = folded gradient of xin0: [encode_x | encode_c_local])
        dim, Fc = m.dim, m.c_local_features
        nvox = X * Y * Z
        gx = g.slice(0, dim)
        bx = torch.zeros(dim, dtype=torch.float32, device=dev)
        dwx = torch.zeros((dim, F), dtype=torch.float32, device=dev)
        call("tdb_cl_nc_outer", gx.ptr, gx.ld, x_in.data_ptr(), F * nvox, dwx.data_ptr(), bx.data_ptr(), B, X, Y, Z, dim, F, dt, s())
        grads["encode_x.weight"] = dwx.reshape(m.encode_x.weight.shape)
        grads["encode_x.bias"] = bx
        g_c_local = None
        if Fc > 0:
            gc = g.slice(dim, dim)
            dwc = torch.zeros((dim, Fc), dtype=torch.float32, device=dev)
            bc = torch.zeros(dim, dtype=torch.float32, device=dev)
            call("tdb_cl_nc_outer", gc.ptr, gc.ld, c_local.data_ptr(), 0, dwc.data_ptr(), bc.data_ptr(), B, X, Y, Z, dim, Fc, dt, s())
            grads["encode_c_local.weight"] = dwc.reshape(m.encode_c_local.weight.shape)
            grads["encode_c_local.bias"] = bc
            wct = m.encode_c_local.weight.detach().reshape(dim, Fc).t().contiguous()  # (Fc, dim)
            zb = torch.zeros(Fc, dtype=torch.float32, device=dev)
            per_sample = torch.empty((B, Fc, X, Y, Z), dtype=torch.float32, device=dev)
            call("tdb_decode_output", gc.ptr, gc.ld, wct.data_ptr(), zb.data_ptr(), per_sample.data_ptr(), B, X, Y, Z, dim, Fc, dt, s())
            g_c_local = per_sample.sum(0)

        # timestep MLP + FiLM projections (ddpm.py:447-452, :184): two launches of tdb_time_film_bwd (no cuBLAS / torch ops, so
        # the whole program captures into a CUDA graph in the default capture mode)
        pc = m.process_c
        w = eng.weights()
        R, dimc = eng.film_rows, m.dim
        g_film_w = torch.empty((R, dimc), dtype=torch.float32, device=dev)
        g_film_b = torch.empty(R, dtype=torch.float32, device=dev)
        g_w1, g_b1 = torch.empty_like(pc[0].weight), torch.empty_like(pc[0].bias)
        g_w2, g_b2 = torch.empty_like(pc[2].weight), torch.empty_like(pc[2].bias)
        dc = torch.empty((B, dimc), dtype=torch.float32, device=dev)
        call("tdb_time_film_bwd", t.data_ptr(), m.encode_t.scale.data_ptr(), m.encode_t.bias.data_ptr(), pc[0].weight.data_ptr(),
             pc[0].bias.data_ptr(), pc[2].weight.data_ptr(), pc[2].bias.data_ptr(), w["film_w"].data_ptr(), p["c"].data_ptr(),
             d_film.data_ptr(), g_film_w.data_ptr(), g_film_b.data_ptr(), g_w1.data_ptr(), g_b1.data_ptr(), g_w2.data_ptr(),
             g_b2.data_ptr(), dc.data_ptr(), B, dimc, R, s())
        grads["process_c.0.weight"], grads["process_c.0.bias"] = g_w1, g_b1
        grads["process_c.2.weight"], grads["process_c.2.bias"] = g_w2, g_b2
        off = 0
        for n in eng.block_order:
            rows = eng.blocks[n].blk.project_onto_scale_shift.weight.shape[0]
            grads[f"{self.prefix[n]}.project_onto_scale_shift.weight"] = g_film_w[off : off + rows]
            grads[f"{self.prefix[n]}.project_onto_scale_shift.bias"] = g_film_b[off : off + rows]
            off += rows
        self._join()
        return grads, g_c_local
