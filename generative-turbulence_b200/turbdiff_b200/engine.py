"""Static launch program for the denoiser: workspace planning + kernel sequencing.

The nn.Module tree in ``models/ddpm.py`` only owns the parameters (same names as the
reference, for checkpoint compatibility).  The compute is organised B200-first: for a given
(batch, grid, precision) a *plan* lays every activation out once in HBM as halo grids
(channels-last, replicate halo materialised, skip tensors written straight into the channel
slice of the consumer's concat buffer), and a flat list of C-ABI kernel launches walks it.
Nothing is allocated while the program runs, so a whole denoise step is CUDA-graph capturable.

Reference semantics implemented here: DenoisingModel.forward ddpm.py:477-505,
UNet.forward :351-372, ResnetBlock :190-197, Block :168-177, Attention :295-308.
"""

from __future__ import annotations

import contextlib
import gc
from dataclasses import dataclass

import os

import torch

from . import _lib
from ._lib import BF16, F32, PW_NOHALO, PW_SILU, call, ptr

GN_EPS = 1e-5


@dataclass
class View:
    """A halo-grid tensor (or a channel slice of one): base tensor, channel offset/count, pitch."""

    t: torch.Tensor  # [B, Xp, Yp, Zp, ld]
    c0: int
    C: int
    level: int

    @property
    def ld(self) -> int:
        return self.t.shape[-1]

    @property
    def ptr(self) -> int:
        return self.t.data_ptr() + self.c0 * self.t.element_size()

    def slice(self, c0: int, C: int) -> "View":
        return View(self.t, self.c0 + c0, C, self.level)


@contextlib.contextmanager
def _gc_paused():
    """No cyclic garbage collection while a CUDA graph is being captured: a collection that happens to run inside the capture
    and releases an OLDER CUDAGraph / its memory pool (cudaGraphExecDestroy, cudaFree) is "an operation not permitted when
    stream is capturing" and invalidates a strict capture (seen in the test suite, which builds many engines per process)."""
    was = gc.isenabled()
    gc.collect()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


def level_sizes(spatial, levels):
    """max(int(s/2), 3) per axis per level (ddpm.py:358)."""
    sizes = [tuple(int(s) for s in spatial)]
    for _ in range(levels):
        sizes.append(tuple(max(int(s * 0.5), 3) for s in sizes[-1]))
    return sizes


class _BlockParams:
    """Flat parameter handles of one ResnetBlock."""

    def __init__(self, blk, film_offset):
        self.blk = blk
        self.cin = blk.block1.conv.in_channels
        self.cout = blk.block1.conv.out_channels
        self.has_proj = not isinstance(blk.conv, torch.nn.Identity)
        self.film_offset = film_offset


class DenoiserEngine:
    def __init__(self, model, precision: str = "fp32"):
        if precision not in ("fp32", "bf16"):
            raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
        self.model = model
        self.precision = precision
        self.dt = F32 if precision == "fp32" else BF16
        self.tdtype = torch.float32 if precision == "fp32" else torch.bfloat16
        self.fused_stats = precision == "bf16"
        self.fold = True
        self.fold2 = True
        self.win = True       # row-window CTA-pair kernel for resident-weight layers with Cout >= 64
        self._level_zp = {}   # level -> Z + 2 of the plan in use (the window kernel needs Z + 2 <= 63)
        self.fold_wide = True
        self.win_center = os.environ.get("TURBDIFF_B200_WIN_CENTER", "1") != "0"  # bottleneck convolutions on the row-window pair kernel
        self.fuse_proj = True
        self.use_graph = True  # p_sample_loop replays the denoiser from a CUDA graph
        self.fuse_tail = os.environ.get("TURBDIFF_B200_FUSE_TAIL", "1") != "0"  # sampler: fused step tail (tdb_step_tail)
        # training: weight repack + forward program and the backward program are replayed from two CUDA graphs
        # (TURBDIFF_B200_TRAIN_GRAPH=0 keeps the eager launch programs)
        self.train_graph = os.environ.get("TURBDIFF_B200_TRAIN_GRAPH", "1") != "0"
        self._train_graphs = {}
        self._train_replay = None
        self.replayed_launches = 0  # C-ABI launches re-issued by training-graph replays (tdb_launch_count() only sees eager ones)
        self._plans = {}
        self._wcache = None
        self._wversion = None
        self._wgen = 0        # bumped whenever _wcache is REPLACED: captured sampler graphs hold its addresses
        self.graph_fallbacks = 0  # CUDA-graph captures that failed and fell back to the eager launch programs
        self._last_train_key = None
        # training: weight-gradient kernels on a side stream, overlapping the bandwidth-bound passes of the next layer
        self.wgrad_side_stream = os.environ.get("TURBDIFF_B200_WGRAD_STREAM", "1") != "0"
        self._side = {}
        self.grad_sync = None  # optional hook(flat_grads): in-place data-parallel reduction of the backward program's flat gradient buffer

        m = model
        un = m.u_net
        self.blocks: dict[str, _BlockParams] = {}
        off = 0
        order = [("decode0", m.decode[0])]
        order += [(f"down{i}", b) for i, b in enumerate(un.downsampling_blocks)]
        order += [(f"up{i}", b) for i, b in enumerate(un.upsampling_blocks)]
        order += [("center0", un.center_block[0]), ("center2", un.center_block[2])]
        for name, blk in order:
            bp = _BlockParams(blk, off)
            self.blocks[name] = bp
            off += 2 * bp.cout
        self.film_rows = off
        self.block_order = [n for n, _ in order]

    def side_stream(self, device, role="wgrad"):
        key = (str(device), role)
        if key not in self._side:
            self._side[key] = torch.cuda.Stream(device=device)
        return self._side[key]

    # ------------------------------------------------------------------ derived weight cache
    def _param_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.model.parameters())

    def _set_geometry(self, spatial):
        """Per-level Z + 2 of the grid about to be processed (fold_kind() consults it for the row-window kernel)."""
        sizes = level_sizes(tuple(spatial), self.model.u_net_levels)
        self._level_zp = {lvl: sz[2] + 2 for lvl, sz in enumerate(sizes)}

    def weights(self, train: bool = False):
        """Kernel-layout copies of the conv weights, refreshed when any parameter changed
        (optimizer step / load_state_dict).  Layout: fp32 [ntaps][Cin][Cout]; bf16 [Cout][ntaps*Cin].
        bf16 path: every layout comes out of ONE launch (tdb_pack_conv_weights_batch, 64 weights per launch); train=True
        also derives the input-gradient layouts the backward program will ask for (_dgrad_weights) in the same call - in
        training all of them are rebuilt every step, and 62 one-weight launches were ~0.7 ms at the head of the step."""
        ver = (self._param_version(), tuple(sorted(self._level_zp.items())))  # kernel choice depends on the grid's Z
        if self._wcache is not None and ver == self._wversion:
            return self._wcache
        m = self.model
        att = m.u_net.center_block[1].fn.fn
        L = m.u_net_levels
        jobs = []  # (key, conv, level) in the order the forward program uses them
        for name in [f"down{i}" for i in range(L)] + ["center0", "attn", "center2"] + [f"up{i}" for i in range(L)] + ["decode0"]:
            if name == "attn":
                jobs.append(("attn.qkv", att.to_qkv, L))
                jobs.append(("attn.out", att.to_out, L))
                continue
            bp, lvl = self.blocks[name], self._block_level(name)
            jobs.append((f"{name}.conv1", bp.blk.block1.conv, lvl))
            jobs.append((f"{name}.conv2", bp.blk.block2.conv, lvl))
            if bp.has_proj:
                jobs.append((f"{name}.proj", bp.blk.conv, lvl))
        assert len(jobs) == 2 * len(self.blocks) + sum(bp.has_proj for bp in self.blocks.values()) + 2
        items = [(conv.weight.detach(), lvl, False) for _, conv, lvl in jobs]
        if train and self.precision == "bf16":
            items += [(conv.weight.detach(), lvl, True) for _, conv, lvl in jobs]
        packed = self.pack_many(items)
        w = {key: packed[i] for i, (key, _, _) in enumerate(jobs)}
        if len(items) > len(jobs):
            w["dgrad"] = {key: packed[len(jobs) + i] for i, (key, _, _) in enumerate(jobs)}
        # stacked over blocks and transposed to (dim, film_rows) for coalesced reads in tdb_time_film
        w["film_w"] = torch.cat([self.blocks[n].blk.project_onto_scale_shift.weight.detach() for n in self.block_order]).t().contiguous().float()
        w["film_b"] = torch.cat([self.blocks[n].blk.project_onto_scale_shift.bias.detach() for n in self.block_order]).contiguous().float()
        self._wcache, self._wversion = w, ver
        self._wgen += 1
        return w

    @staticmethod
    def pad_rows(size) -> int:
        X, Y, Z = size
        return (Y + 2) * (Z + 2) + 2 * (Z + 2) + 256

    def fold_kind(self, ntaps, cin, cout, level):
        """Which bf16 kernel a 3x3x3 convolution runs on (None = per-tap kernel tdb_conv3d_bf16):
        "winz"  kz-folded row-window CTA pair, resident weights: Cout 64 and 128 -> 32 (narrow, full resolution);
        "win"   row-window CTA pair: 32 -> 128 resident, and every Cout % 128 == 0 layer as N tiles with streamed weights;
        "fold2" kz-folded CTA pair with nine tiles per chunk: what the row-window kernels cannot take (256 -> 64);
        "fold"  single-CTA kz-folded kernel: Cout <= 64 leftovers (32 -> 32 switches to the paired-row kernel in _conv
                when the input pitch is exactly 32, 128-byte aligned and Z + 2 is even).
        The bottleneck level (tiny M, huge K) stays on the per-tap kernel with split-K."""
        if self.precision != "bf16" or not self.fold or ntaps != 27:
            return None
        zp = self._level_zp.get(level, 1 << 30)
        resident = cin % 32 == 0 and 27 * cin * cout <= 116 * 1024
        if self.win and resident and zp <= 63 and ((cout == 32 and cin >= 64) or cout == 64):
            # row-window CTA pair with kz folded into N (N = 3*Cout): one window per kx instead of nine tiles per chunk
            # (128->32: 0.48 -> 0.35 ms at B=4, L2 -> SM traffic / 1.6), fat MMAs for the narrow layers
            return "winz"
        if self.win and resident and zp <= 63 and cout == 128:
            # row-window CTA-pair kernel, N = Cout: weights resident, activations staged 3x instead of 9x
            return "win"
        if cout in (16, 32, 64):
            pair = self.fold2 and cin % 64 == 0 and cout in (32, 64) and 9 * cin * 3 * cout * 2 > 112 * 1024
            return "fold2" if pair else "fold"
        if self.fold2 and self.fold_wide and self.win and zp <= 63 and cout % 512 == 0 and cout > 512 and cin % 64 == 0:
            # Cout > 512 (the 256 -> 1024 input gradient of up0.block1): the row-window kernel in launches of 512 channels
            # (the per-tap layout's row blocks are contiguous); _conv walks the halves
            return "win"
        if self.fold2 and self.fold_wide and cout % 128 == 0 and cout <= 512 and cin % 64 == 0:
            # wide layers: N tiles of 128 channels with streamed weights; the row-window kernel double-buffers its
            # accumulators and keeps all 128 rows of a tile (3-20 % faster than the kz-folded pair kernel here).
            # The bottleneck level runs on it too (round 2): with 2800 haloed rows at B = 8 the per-tap split-K kernel
            # pushed every weight byte through the L2 -> SM fabric 22 times (one 128-row M tile each) and needed a
            # finalize and a moments pass (48 + 6 + 9 us per convolution); 256-row pair tiles with fused moments do not.
            if level == self.model.u_net_levels and not self.win_center:
                return None
            return "win" if (self.win and zp <= 63) else "fold2"
        return None

    def _pack_spec(self, wt, level, dgrad):
        """(Cout, Cin, taps, folded, N tile, shape of the packed matrix) of a (Cout, Cin, k, k, k) weight."""
        cout, cin = wt.shape[:2]
        taps = wt.shape[2] * wt.shape[3] * wt.shape[4]
        lo, li = (cin, cout) if dgrad else (cout, cin)  # logical output / input channels of the packed matrix
        folded = self.precision == "bf16" and self.fold_kind(taps, li, lo, level) not in (None, "win")
        tile = lo if lo < 128 else 128
        return cout, cin, taps, folded, tile, ((3 * lo, 9 * li) if folded else (lo, taps * li))

    def pack_many(self, items):
        """pack_conv of every (weight, level, dgrad) item.  bf16 path on CUDA: one flat bf16 buffer and one launch per 64
        weights (tdb_pack_conv_weights_batch); anything else item by item."""
        ok = self.precision == "bf16" and len(items) > 0 and all(
            wt.is_cuda and wt.dtype == torch.float32 and wt.shape[2] * wt.shape[3] * wt.shape[4] in (1, 27) for wt, _, _ in items)
        if not ok:
            return [self.pack_conv(wt, lvl, dg) for wt, lvl, dg in items]
        specs = [self._pack_spec(wt, lvl, dg) for wt, lvl, dg in items]
        offs, total = [], 0
        for sp in specs:
            offs.append(total)
            total += -(-(sp[5][0] * sp[5][1]) // 256) * 256  # every layout starts on a 512-byte boundary (TMA base addresses)
        flat = torch.empty(total, dtype=torch.bfloat16, device=items[0][0].device)
        srcs = [wt.detach().contiguous() for wt, _, _ in items]
        outs = [flat[o : o + sp[5][0] * sp[5][1]].view(sp[5]) for o, sp in zip(offs, specs)]
        table = (_lib.PackJob * len(items))()
        for j, (src, dst, sp, (_, _, dg)) in enumerate(zip(srcs, outs, specs, items)):
            table[j] = _lib.PackJob(src.data_ptr(), dst.data_ptr(), sp[0], sp[1], sp[2], 1 if sp[3] else 0, sp[4], 1 if dg else 0)
        import ctypes

        call("tdb_pack_conv_weights_batch", ctypes.cast(table, ctypes.c_void_p), len(items), _lib.stream_ptr())
        return outs

    def pack_conv(self, wt, level, dgrad: bool = False):
        """Kernel layout of a (Cout, Cin, k, k, k) weight for the kernel fold_kind() selects at `level`.  dgrad=True: the
        weights of the input-gradient convolution, W'[ci][co][k] = W[co][ci][2-k] (logical Cout' = Cin, Cin' = Cout).
        bf16 path on CUDA: one launch of tdb_pack_conv_weights per weight, straight from the fp32 parameter (as torch ops
        this was a strided permute copy, a flip and a cast per weight: 178 launches per training step)."""
        cout, cin, taps, folded, tile, shape = self._pack_spec(wt, level, dgrad)
        lo, li = (cin, cout) if dgrad else (cout, cin)
        if self.precision == "bf16" and wt.is_cuda and wt.dtype == torch.float32 and taps in (1, 27):
            src = wt.detach().contiguous()
            dst = torch.empty(shape, dtype=torch.bfloat16, device=wt.device)
            call("tdb_pack_conv_weights", src.data_ptr(), dst.data_ptr(), cout, cin, taps, 1 if folded else 0, tile, 1 if dgrad else 0,
                 _lib.stream_ptr())
            return dst
        if dgrad:
            wt = wt.flip(2, 3, 4).transpose(0, 1)
            cout, cin = lo, li
        if self.precision == "fp32":
            return wt.permute(2, 3, 4, 1, 0).reshape(taps, cin, cout).contiguous().float()
        if folded:
            # kz folded into N, N tiles of <= 128 channels: row = (tile*3 + kz)*T + co, col = (kx*3+ky)*Cin + ci
            return (wt.reshape(cout // tile, tile, cin, 3, 3, 3).permute(0, 5, 1, 3, 4, 2)
                    .reshape(3 * cout, 9 * cin).contiguous().to(torch.bfloat16))
        return wt.permute(0, 2, 3, 4, 1).reshape(cout, taps * cin).contiguous().to(torch.bfloat16)

    # ------------------------------------------------------------------ workspace
    def plan(self, B, spatial, device):
        key = (B, tuple(spatial), str(device))
        if key in self._plans:
            return self._plans[key]
        m = self.model
        L = m.u_net_levels
        dim = m.dim
        sizes = level_sizes(spatial, L)
        td = self.tdtype

        def grid(level, C):
            # every halo grid is allocated with `pad` zero rows in front and behind: the folded
            # convolution reads its row-shifted tiles through an overlapping TMA view (no OOB fill)
            X, Y, Z = sizes[level]
            rows, pad = B * (X + 2) * (Y + 2) * (Z + 2), self.pad_rows(sizes[level])
            flat = torch.zeros((rows + 2 * pad, C), dtype=td, device=device)
            return View(flat[pad : pad + rows].view(B, X + 2, Y + 2, Z + 2, C), 0, C, level)

        p = {"sizes": sizes, "B": B}
        c_in0 = dim + (dim if m.c_local_features > 0 else 0)
        p["xin0"] = grid(0, c_in0)
        cd = [dim * 2 ** (l + 1) for l in range(L)]  # channels of down block l's output / skip
        p["cat"] = [grid(l, 2 * cd[l]) for l in range(L)]
        p["xin"] = [p["xin0"]] + [grid(l, cd[l - 1]) for l in range(1, L)]
        center = dim * 2**L
        p["center_in"] = grid(L, center)
        p["center0"] = grid(L, center)
        p["center1"] = grid(L, center)
        p["center2"] = grid(L, center)
        p["up_out"] = [grid(l, dim * 2**l) for l in range(L)]
        p["dec_out"] = grid(0, dim)
        att = m.u_net.center_block[1].fn.fn
        hid = att.heads * att.dim_head
        p["attn_norm"] = grid(L, center)
        p["attn_qkv"] = grid(L, 3 * hid)
        p["attn_o"] = grid(L, hid)
        p["attn_proj"] = grid(L, center)
        # scratch per (level, channel count) with the EXACT channel pitch: a 32-channel tensor kept in a 64-channel-pitch
        # buffer uses half of every 128-byte line, and DRAM/L2 move whole lines - ncu showed 2x the algorithmic read
        # traffic for the 32->32 convolutions (and every streaming kernel over those tensors) with shared buffers
        shapes = sorted({(self._block_level(name), bp.cout) for name, bp in self.blocks.items()})
        max_c = {}
        for lvl, c in shapes:
            max_c[lvl] = max(max_c.get(lvl, 0), c)
        p["raw"] = {k: grid(*k) for k in shapes}
        p["act"] = {k: grid(*k) for k in shapes}
        p["res"] = {k: grid(*k) for k in shapes}
        # fp32 split-K workspace of the tensor-core convolution: large enough for the two deepest levels
        deep = max(1, L - 1)
        Xd, Yd, Zd = sizes[deep]
        p["splitk"] = torch.zeros(B * (Xd + 2) * (Yd + 2) * (Zd + 2) * 1024, dtype=torch.float32, device=device)
        p["grid"] = grid  # allocator for the (lazily built) training buffers
        p["max_c"] = max_c
        n_norms = 2 * len(self.blocks) + 1
        gmax = max(self._groups(bp.cout) for bp in self.blocks.values())
        # one flat slot per norm layer, used as [B][G][2] doubles (sum, sum of squares)
        p["stats"] = torch.zeros((n_norms, B * gmax * 2), dtype=torch.float64, device=device)
        p["film"] = torch.zeros((B, self.film_rows), dtype=torch.float32, device=device)
        p["c"] = torch.zeros((B, dim), dtype=torch.float32, device=device)
        p["eps"] = torch.zeros((B, m.out_features, *spatial), dtype=torch.float32, device=device)
        self._plans[key] = p
        return p

    def _block_level(self, name):
        L = self.model.u_net_levels
        if name == "decode0":
            return 0
        if name.startswith("down"):
            return int(name[4:])
        if name.startswith("up"):
            return L - 1 - int(name[2:])
        return L

    def _groups(self, C):
        g = self.model.norm_groups
        return C if g is None else g

    # ------------------------------------------------------------------ kernels
    def can_add1x1(self, cin, cout, level) -> bool:
        """out = conv3x3x3(x) + conv1x1(x2) in one launch: the row-window pair kernel without kz folding."""
        return self.precision == "bf16" and self.fold_kind(27, cin, cout, level) == "win"

    def _conv(self, p, x: View, w, bias, out: View, ntaps, stats=None, G=0, all_rows=False, proj=None, add1x1=None):
        """3x3x3 / 1x1x1 convolution over halo grids.  all_rows: also store the halo rows of the output
        (input-gradient convolutions; the fp32 kernel always stores every row).  add1x1 = (x2 view with x.C channels,
        w2 [Cout][Cin] bf16): the 1x1 convolution of a second input is accumulated in the same tile (can_add1x1())."""
        B = p["B"]
        X, Y, Z = p["sizes"][x.level]
        s = _lib.stream_ptr()
        flags = _lib.CONV_ALL_ROWS if all_rows else 0
        if self.precision == "bf16" and ntaps == 27 and out.C > 512 and self.fold_kind(ntaps, x.C, out.C, x.level) == "win":
            # more than 512 output channels: one launch per block of 512 (rows h*512 .. of the [Cout][27*Cin] weights)
            assert stats is None and proj is None and out.C % 512 == 0
            for h in range(out.C // 512):
                wh = w[h * 512 : (h + 1) * 512]
                bh = None if bias is None else bias[h * 512 : (h + 1) * 512]
                a2 = None if add1x1 is None else (add1x1[0], add1x1[1][h * 512 : (h + 1) * 512])
                self._conv(p, x, wh, bh, out.slice(h * 512, 512), ntaps, all_rows=all_rows, add1x1=a2)
            return
        if add1x1 is not None:
            x2, w2 = add1x1
            assert self.can_add1x1(x.C, out.C, x.level) and x2.C == x.C and stats is None and proj is None
            call("tdb_conv3d_bf16_win_add1x1", x.ptr, x.ld, w.data_ptr(), ptr(bias), out.ptr, out.ld, B, X, Y, Z, x.C, out.C, flags,
                 x2.ptr, x2.ld, w2.data_ptr(), s)
        elif self.precision == "fp32":
            call("tdb_conv3d_f32", x.ptr, x.ld, w.data_ptr(), ptr(bias), out.ptr, out.ld, B, X, Y, Z, x.C, out.C, ntaps, s)
        elif self.fold_kind(ntaps, x.C, out.C, x.level) in ("win", "winz"):
            pw, pb, pv = proj if proj is not None else (None, None, None)
            call("tdb_conv3d_bf16_" + self.fold_kind(ntaps, x.C, out.C, x.level), x.ptr, x.ld, w.data_ptr(), ptr(bias), out.ptr, out.ld, B, X, Y, Z, x.C, out.C, ptr(stats), G,
                 flags, ptr(pw), ptr(pb), pv.ptr if pv is not None else None, pv.ld if pv is not None else 0, s)
        elif self.fold_kind(ntaps, x.C, out.C, x.level) == "fold2":
            # proj = (weights [Cout][Cin] bf16, bias, output view): the block's 1x1 residual projection, fused
            pw, pb, pv = proj if proj is not None else (None, None, None)
            call("tdb_conv3d_bf16_fold2", x.ptr, x.ld, self.pad_rows((X, Y, Z)), w.data_ptr(), ptr(bias), out.ptr, out.ld,
                 B, X, Y, Z, x.C, out.C, ptr(stats), G, flags, ptr(pw), ptr(pb), pv.ptr if pv is not None else None,
                 pv.ld if pv is not None else 0, s)
        elif (self.fold_kind(ntaps, x.C, out.C, x.level) == "fold" and self.win and x.C == 32 and out.C == 32 and x.ld == 32
              and x.ptr % 128 == 0 and (Z + 2) % 2 == 0 and Z + 2 <= 128):
            # 32 -> 32 with an exact input pitch: two grid rows per 128-byte TMA row (same folded weight layout)
            call("tdb_conv3d_bf16_winp", x.ptr, x.ld, w.data_ptr(), ptr(bias), out.ptr, out.ld, B, X, Y, Z, x.C, out.C, ptr(stats), G,
                 flags, s)
        elif self.fold_kind(ntaps, x.C, out.C, x.level) == "fold":
            call("tdb_conv3d_bf16_fold", x.ptr, x.ld, self.pad_rows((X, Y, Z)), w.data_ptr(), ptr(bias), out.ptr, out.ld,
                 B, X, Y, Z, x.C, out.C, ptr(stats), G, flags, s)
        else:
            rows = B * (X + 2) * (Y + 2) * (Z + 2)
            scratch = p["splitk"] if rows * out.C <= p["splitk"].numel() else None
            call("tdb_conv3d_bf16", x.ptr, x.ld, w.data_ptr(), ptr(bias), out.ptr, out.ld, B, X, Y, Z, x.C, out.C, ntaps,
                 ptr(stats), G, flags, ptr(scratch), s)

    def _stats(self, p, x: View, stats, G):
        X, Y, Z = p["sizes"][x.level]
        call("tdb_gn_stats", x.ptr, x.ld, stats.data_ptr(), p["B"], X, Y, Z, x.C, G, self.dt, _lib.stream_ptr())

    def _pointwise(self, p, raw: View, stats, norm, film_ptr, res: View | None, out: View, flags, G=1):
        X, Y, Z = p["sizes"][raw.level]
        call(
            "tdb_pointwise", raw.ptr, raw.ld, ptr(stats), ptr(norm.weight) if norm is not None else None,
            ptr(norm.bias) if norm is not None else None, film_ptr, self.film_rows,
            res.ptr if res is not None else None, res.ld if res is not None else 0, out.ptr, out.ld,
            p["B"], X, Y, Z, raw.C, G, GN_EPS, flags, self.dt, _lib.stream_ptr(),
        )

    def _trilinear(self, p, x: View, out: View):
        Xi, Yi, Zi = p["sizes"][x.level]
        Xo, Yo, Zo = p["sizes"][out.level]
        call("tdb_trilinear", x.ptr, x.ld, Xi, Yi, Zi, out.ptr, out.ld, Xo, Yo, Zo, p["B"], x.C, self.dt, _lib.stream_ptr())

    def can_fuse_proj(self, x: View, cout) -> bool:
        """The 1x1 residual projection rides on conv1's centre-tap tiles when conv1 runs on a CTA pair: Cout <= 64 on the
        kz-folded / resident row-window kernels, and any 128-channel N tiling of the streamed row-window kernel."""
        if not self.fuse_proj:
            return False
        kind = self.fold_kind(27, x.C, cout, x.level)
        # (more than 512 output channels run as several launches of the row-window kernel: no fused projection there)
        return (cout <= 64 and kind in ("fold2", "win", "winz")) or (kind == "win" and cout % 128 == 0 and cout <= 512)

    def _norm_conv(self, p, x: View, w, conv, norm, raw: View, stats_slot, proj=None):
        """conv (+bias) followed by GroupNorm moments of its output."""
        G = self._groups(raw.C)
        stats = p["stats"][stats_slot]
        cpg = raw.C // G
        if self.fold_kind(27, x.C, raw.C, x.level) is not None:
            # (Cout > 512 is split into launches of 512 channels, whose [B][G] moment slots would not line up: gn_stats pass)
            fused = self.fused_stats and cpg % 2 == 0 and raw.C <= 512
        else:
            fused = self.fused_stats and (cpg % 16 == 0 or 16 % cpg == 0)
        self._conv(p, x, w, conv.bias, raw, 27, stats if fused else None, G, proj=proj)
        if not fused:
            self._stats(p, raw, stats, G)
        return stats, G

    def _saved(self, p, name):
        """Per-block activation buffers kept for the backward program (training mode only)."""
        sv = p.setdefault("saved", {})
        if name not in sv:
            bp = self.blocks[name]
            lvl = self._block_level(name)
            sv[name] = {k: p["grid"](lvl, bp.cout) for k in ("raw1", "act1", "raw2")}
        return sv[name]

    def _resblock(self, p, name, x: View, out: View, slot, train=False, defer_out=False):
        """defer_out: stop before the block's last pointwise (GroupNorm + SiLU + residual) and leave its operands in
        p["tail"] - the sampler's fused step tail (tdb_step_tail) consumes them."""
        bp = self.blocks[name]
        w = self.weights()
        lvl = x.level
        blk = bp.blk
        if train:
            sv = self._saved(p, name)
            raw, act, raw_b = sv["raw1"], sv["act1"], sv["raw2"]
            sv["x"], sv["out"], sv["slot"] = x, out, slot
        else:
            raw = raw_b = p["raw"][(lvl, bp.cout)]
            act = p["act"][(lvl, bp.cout)]
        film_ptr = p["film"].data_ptr() + 4 * bp.film_offset
        proj = None
        if bp.has_proj and self.can_fuse_proj(x, bp.cout):
            proj = (w[f"{name}.proj"], blk.conv.bias, p["res"][(lvl, bp.cout)])
        st, G = self._norm_conv(p, x, w[f"{name}.conv1"], blk.block1.conv, blk.block1.norm, raw, slot, proj=proj)
        self._pointwise(p, raw, st, blk.block1.norm, film_ptr, None, act, PW_SILU, G)
        raw = raw_b
        st, G = self._norm_conv(p, act, w[f"{name}.conv2"], blk.block2.conv, blk.block2.norm, raw, slot + 1)
        if bp.has_proj:
            res = p["res"][(lvl, bp.cout)]
            if proj is None:
                self._conv(p, x, w[f"{name}.proj"], blk.conv.bias, res, 1)
        else:
            res = x
        if defer_out:
            p["tail"] = {"raw": raw, "stats": st, "G": G, "norm": blk.block2.norm, "res": res}
            return
        self._pointwise(p, raw, st, blk.block2.norm, None, res, out, PW_SILU, G)

    def _attention(self, p, x: View, out: View, slot):
        m = self.model
        w = self.weights()
        pre = m.u_net.center_block[1].fn
        att = pre.fn
        X, Y, Z = p["sizes"][x.level]
        G = self._groups(x.C)
        stats = p["stats"][slot]
        self._stats(p, x, stats, G)
        hn = p["attn_norm"]
        self._pointwise(p, x, stats, pre.norm, None, None, hn, PW_NOHALO, G)
        self._conv(p, hn, w["attn.qkv"], None, p["attn_qkv"], 1)
        call("tdb_attention", p["attn_qkv"].ptr, p["attn_qkv"].ld, p["attn_o"].ptr, p["attn_o"].ld, p["B"], X, Y, Z,
             att.heads, att.dim_head, self.dt, _lib.stream_ptr())
        self._conv(p, p["attn_o"], w["attn.out"], att.to_out.bias, p["attn_proj"], 1)
        self._pointwise(p, p["attn_proj"], None, None, None, x, out, 0)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x: torch.Tensor, t: torch.Tensor, c_local: torch.Tensor | None, taps: dict | None = None, train: bool = False,
                c_static: bool = False, tail: bool = False, encode_x: bool = True):
        """eps = U-Net(x, t, c_local).  x (B,F,X,Y,Z) fp32 CUDA contiguous, t int64 (B,).
        train=True keeps every block's intermediates in dedicated buffers for `backward`.
        Sampler mode: tail=True stops after decode.0's last convolution (the fused step tail finishes the step and
        returns nothing here); encode_x=False leaves the x half of the level-0 input buffer as the previous step's
        tail wrote it."""
        m = self.model
        _lib.require_cuda(x, "x")
        if x.dtype != torch.float32:
            raise RuntimeError("turbdiff_b200: x must be float32 at the module boundary")
        x = x.contiguous()
        t = t.to(device=x.device, dtype=torch.int64).contiguous()
        B, F = x.shape[:2]
        spatial = tuple(x.shape[2:])
        L = m.u_net_levels
        p = self.plan(B, spatial, x.device)
        self._set_geometry(spatial)
        w = self.weights(train=train)
        s = _lib.stream_ptr()
        X, Y, Z = spatial
        Fc = m.c_local_features
        if Fc > 0:
            if c_local is None:
                raise RuntimeError("turbdiff_b200: model expects local conditioning but C has none")
            c_local = c_local.to(torch.float32).contiguous()
            if tuple(c_local.shape) != (Fc, X, Y, Z):
                raise RuntimeError(f"c_local shape {tuple(c_local.shape)} != {(Fc, X, Y, Z)}")

        p["stats"].zero_()
        pc = m.process_c
        call("tdb_time_film", t.data_ptr(), m.encode_t.scale.data_ptr(), m.encode_t.bias.data_ptr(), pc[0].weight.data_ptr(),
             pc[0].bias.data_ptr(), pc[2].weight.data_ptr(), pc[2].bias.data_ptr(), w["film_w"].data_ptr(),
             w["film_b"].data_ptr(), p["c"].data_ptr(), p["film"].data_ptr(), B, m.dim, self.film_rows, s)
        xin0 = p["xin0"]
        # encode_c_local(c_local) does not depend on the step (the reference recomputes it, ddpm.py:480 TODO):
        # its half of the concat buffer is rewritten only when c_local or the encoder weights changed
        # (c_static=True is the caller's promise that c_local and the weights are those of the previous call on this plan)
        parts = 1 if (c_static and Fc > 0 and p.get("c_valid") and not train) else 3
        if not encode_x:
            parts &= 2
        p["c_valid"] = Fc > 0
        if parts:
            call("tdb_encode_input", x.data_ptr(), ptr(c_local), m.encode_x.weight.data_ptr(), m.encode_x.bias.data_ptr(),
                 ptr(m.encode_c_local.weight) if Fc > 0 else None, ptr(m.encode_c_local.bias) if Fc > 0 else None,
                 xin0.ptr, xin0.ld, B, F, Fc, m.dim, X, Y, Z, parts, self.dt, s)

        def tap(name, v: View):
            if taps is not None:
                taps[name] = self.to_ncdhw(v)

        slot = 0
        cur = xin0
        for l in range(L):
            cd = p["cat"][l].C // 2
            out = p["cat"][l].slice(cd, cd)
            self._resblock(p, f"down{l}", cur, out, slot, train)
            slot += 2
            tap(f"down{l}", out)
            nxt = p["xin"][l + 1] if l + 1 < L else p["center_in"]
            self._trilinear(p, out, nxt)
            cur = nxt
        self._resblock(p, "center0", cur, p["center0"], slot, train)
        slot += 2
        tap("center0", p["center0"])
        self._attention(p, p["center0"], p["center1"], slot)
        p["attn_slot"] = slot
        slot += 1
        tap("center1", p["center1"])
        self._resblock(p, "center2", p["center1"], p["center2"], slot, train)
        slot += 2
        tap("center2", p["center2"])
        cur = p["center2"]
        for i in range(L):
            l = L - 1 - i
            cat = p["cat"][l]
            self._trilinear(p, cur, cat.slice(0, cat.C // 2))
            self._resblock(p, f"up{i}", cat, p["up_out"][l], slot, train)
            slot += 2
            tap(f"up{i}", p["up_out"][l])
            cur = p["up_out"][l]
        self._resblock(p, "decode0", cur, p["dec_out"], slot, train, defer_out=tail)
        if tail:
            return None
        tap("decode0", p["dec_out"])
        dec = m.decode[1]
        eps = p["eps"]
        call("tdb_decode_output", p["dec_out"].ptr, p["dec_out"].ld, dec.weight.data_ptr(), dec.bias.data_ptr(), eps.data_ptr(),
             B, X, Y, Z, m.dim, m.out_features, self.dt, s)
        if train:
            p["last_input"] = (x, t, c_local)
            self._last_train_key = (B, tuple(spatial), str(x.device))
        return eps

    # ------------------------------------------------------------------ sampling state / CUDA graph
    def sampler_state(self, B, spatial, device, c_local):
        """Persistent per-plan state of an ancestral-sampling chain: the state tensor x_t, the timestep tensors and a
        copy of c_local, all at fixed addresses, plus a CUDA graph of the denoiser launch program over them.  The
        graph is (re)captured when the kernel-layout weights changed; replays cost one launch instead of ~80."""
        p = self.plan(B, spatial, device)
        m = self.model
        st = p.get("sampler")
        if st is None:
            st = p["sampler"] = {
                "x_t": torch.zeros((B, m.in_features, *spatial), dtype=torch.float32, device=device),
                "x_t2": torch.zeros((B, m.in_features, *spatial), dtype=torch.float32, device=device),  # fused tail: double-buffered state
                "t_vec": torch.zeros(B, dtype=torch.int64, device=device),
                "t_dev": torch.zeros(1, dtype=torch.int32, device=device),
                "c_local": None if c_local is None else torch.zeros_like(c_local, dtype=torch.float32),
                "graph": None, "wver": None, "graph_tail": None, "wver_tail": None,
            }
        if c_local is not None:
            st["c_local"].copy_(c_local)
        # the reference re-encodes C on every call (ddpm.py:496-501); the captured graph only rewrites the x half of the
        # concat buffer, so the c_local half is re-encoded once per chain (forward_graphed) - a training forward or another
        # chain with a different geometry may have overwritten it since
        st["c_dirty"] = True
        return st

    def _encode_c_half(self, st):
        """encode_c_local(c_local) into its channel half of the level-0 concat buffer (one launch, no x half)."""
        m = self.model
        Fc = m.c_local_features
        x_t = st["x_t"]
        if Fc > 0:
            B, F = x_t.shape[:2]
            X, Y, Z = x_t.shape[2:]
            p = self.plan(B, (X, Y, Z), x_t.device)
            xin0 = p["xin0"]
            call("tdb_encode_input", x_t.data_ptr(), ptr(st["c_local"]), m.encode_x.weight.data_ptr(), m.encode_x.bias.data_ptr(),
                 ptr(m.encode_c_local.weight), ptr(m.encode_c_local.bias), xin0.ptr, xin0.ld, B, F, Fc, m.dim, X, Y, Z, 2, self.dt,
                 _lib.stream_ptr())
            p["c_valid"] = True
        st["c_dirty"] = False

    def can_fuse_tail(self) -> bool:
        """tdb_step_tail covers the whole step tail when the model fits its register plan."""
        m = self.model
        return m.dim in (8, 16, 32, 64) and m.in_features <= 4 and m.in_features <= m.out_features <= 8 and self.fuse_tail

    def encode_state(self, st, x: torch.Tensor):
        """encode_x(x) into the x half of the level-0 input buffer (start of a chain in fused-tail mode)."""
        m = self.model
        B, F = x.shape[:2]
        X, Y, Z = x.shape[2:]
        p = self.plan(B, (X, Y, Z), x.device)
        xin0 = p["xin0"]
        Fc = m.c_local_features
        call("tdb_encode_input", x.data_ptr(), ptr(st["c_local"]), m.encode_x.weight.data_ptr(), m.encode_x.bias.data_ptr(),
             ptr(m.encode_c_local.weight) if Fc > 0 else None, ptr(m.encode_c_local.bias) if Fc > 0 else None,
             xin0.ptr, xin0.ld, B, F, Fc, m.dim, X, Y, Z, 1, self.dt, _lib.stream_ptr())

    def step_tail(self, st, x_in, x_out, z, z_bc, x_bcs, mask, coef, t_dev, flags, eps_out=None):
        """The fused tail of one sampling step over the operands forward(tail=True) left in the plan: decode.0's last
        pointwise + decode.1 + posterior update (x_in -> x_out) + encode_x of the next step."""
        m = self.model
        B, F = x_in.shape[:2]
        X, Y, Z = x_in.shape[2:]
        p = self.plan(B, (X, Y, Z), x_in.device)
        tl, xin0, dec = p["tail"], p["xin0"], m.decode[1]
        call("tdb_step_tail", tl["raw"].ptr, tl["raw"].ld, tl["stats"].data_ptr(), tl["norm"].weight.data_ptr(), tl["norm"].bias.data_ptr(),
             tl["res"].ptr, tl["res"].ld, dec.weight.data_ptr(), dec.bias.data_ptr(), m.out_features, x_in.data_ptr(), ptr(z), ptr(z_bc),
             ptr(x_bcs), mask.data_ptr(), coef.data_ptr(), t_dev.data_ptr(), x_out.data_ptr(), ptr(eps_out), F, flags,
             m.encode_x.weight.data_ptr(), m.encode_x.bias.data_ptr(), xin0.ptr, xin0.ld, B, X, Y, Z, m.dim, tl["G"], GN_EPS, self.dt,
             _lib.stream_ptr())

    def forward_graphed(self, st, tail: bool = False):
        """eps for the sampler state `st` (x_t, t_vec, c_local buffers); captures the graph on first use.  tail=True:
        the sampler's fused-tail mode - the program neither encodes x (the previous tail did) nor finishes decode.0 /
        decode.1 (the next tail will); returns None."""
        self._set_geometry(st["x_t"].shape[2:])
        self.weights()
        kw = {"tail": True, "encode_x": False} if tail else {}
        gk, wk = ("graph_tail", "wver_tail") if tail else ("graph", "wver")
        if not self.use_graph:
            static = not st.get("c_dirty", True)
            st["c_dirty"] = False
            return self.forward(st["x_t"], st["t_vec"], st["c_local"], c_static=static, **kw)
        if st[gk] is None or st[wk] != self._wgen:
            # eager warm-up on a side stream (also (re)writes the c_local half), then capture
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.forward(st["x_t"], st["t_vec"], st["c_local"], c_static=False, **kw)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            try:
                with _gc_paused(), torch.cuda.graph(g):
                    self.forward(st["x_t"], st["t_vec"], st["c_local"], c_static=True, **kw)
            except RuntimeError as e:  # not fatal: the same launch program runs eagerly
                import warnings

                warnings.warn(f"turbdiff_b200: CUDA-graph capture of the denoiser failed ({str(e)[:200]}); using the eager launch program")
                self.use_graph = False
                self.graph_fallbacks += 1
                torch.cuda.synchronize()
                st["c_dirty"] = False
                return self.forward(st["x_t"], st["t_vec"], st["c_local"], c_static=False, **kw)
            st[gk], st[wk] = g, self._wgen
            st["c_dirty"] = False  # the warm-up wrote the c_local half from st["c_local"]
        if st.get("c_dirty", True):
            self._encode_c_half(st)
        st[gk].replay()
        if tail:
            return None
        B = st["x_t"].shape[0]
        return self.plan(B, tuple(st["x_t"].shape[2:]), st["x_t"].device)["eps"]

    def backward(self, g_eps: torch.Tensor):
        """Gradients of every parameter (and of c_local) for the last `forward(train=True)`."""
        from .backward import BackwardProgram

        return BackwardProgram(self).run(g_eps)

    # ------------------------------------------------------------------ training step as two CUDA graphs
    def train_forward(self, x, t, c_local):
        """forward(train=True).  With train_graph the kernel-layout weight refresh and the launch program run as ONE
        graph replay over static input buffers (a training step is ~370 kernel launches plus ~1000 small torch ops on
        the host otherwise); the matching backward graph is captured at the same time."""
        if not self.train_graph or _lib.PROFILE is not None or not x.is_cuda:
            self._train_replay = None
            return self.forward(x, t, c_local, train=True)
        key = (x.shape[0], tuple(x.shape[2:]), str(x.device), c_local is None)
        sig = tuple(q.data_ptr() for q in self.model.parameters())
        tg = self._train_graphs.get(key)
        if tg is None or tg["sig"] != sig:
            tg = None
            # strict capture first (thread_local: nothing in the programs needs more); if something outside them - e.g. the
            # garbage collector releasing an older graph's memory on this thread in the middle of the capture - invalidates
            # it, once more in relaxed mode; only then the eager launch programs
            for mode in ("thread_local", "relaxed"):
                try:
                    tg = self._train_graphs[key] = self._capture_train(x, t, c_local, sig, mode)
                    break
                except RuntimeError as e:
                    err = e
                    torch.cuda.synchronize()
            if tg is None:  # a failed capture is not fatal: the same kernels run from the eager launch programs
                import warnings

                warnings.warn(f"turbdiff_b200: CUDA-graph capture of the training step failed ({str(err)[:200]}); using the eager launch programs")
                self.train_graph = False
                self.graph_fallbacks += 1
                self._train_replay = None
                return self.forward(x, t, c_local, train=True)
        tg["x"].copy_(x)
        tg["t"].copy_(t)
        if c_local is not None:
            tg["c"].copy_(c_local)
        tg["fwd"].replay()
        self.replayed_launches += tg["n_fwd"]
        self._last_train_key = key[:3]
        self._train_replay = tg
        return tg["eps"]

    def train_backward(self, g_eps):
        """Backward of the last train_forward: (parameter gradients, gradient of c_local).  Eager mode returns the
        gradients as a dict by parameter name; graph mode returns ONE flat fp32 buffer holding them in the order
        self._train_replay["order"] (phase-1 parameters first, see backward.grad_phase; ["sizes"] / ["shapes"] describe the
        split).  With a data-parallel hook attached (self.grad_sync) the exchange of the phase-1 gradients is started
        between the two backward graphs, so it travels while the down path is still being differentiated; the caller
        finishes it with grad_sync.finish()."""
        tg = self._train_replay
        if tg is None:
            return self.backward(g_eps)
        tg["g_eps"].copy_(g_eps)
        sync = self.grad_sync
        tg["bwd"].replay()
        if sync is not None:
            sync.start(tg["flat"][: tg["n1"]])
        tg["bwd2"].replay()
        if sync is not None:
            sync.start(tg["flat"][tg["n1"] :])
        self.replayed_launches += tg["n_bwd"]
        return tg["flat"], tg["g_c_local"]

    def _capture_train(self, x, t, c_local, sig, capture_mode="thread_local"):
        from .backward import BackwardProgram

        xs = x.detach().to(torch.float32).clone()
        ts = t.detach().to(device=x.device, dtype=torch.int64).clone()
        cs = None if c_local is None else c_local.detach().to(torch.float32).clone()
        gs = torch.zeros((x.shape[0], self.model.out_features, *x.shape[2:]), dtype=torch.float32, device=x.device)
        # eager warm-up on a side stream: plan buffers, kernel attributes and lazy caches exist before the capture
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.forward(xs, ts, cs, train=True)
            BackwardProgram(self).run(gs)
        torch.cuda.current_stream().wait_stream(side)
        pool = torch.cuda.graph_pool_handle()
        g_f, g_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        # capture_error_mode="thread_local": both programs consist of this library's launches plus allocator-backed torch
        # element-wise ops only (the timestep-MLP backward is tdb_time_film_bwd since round 2: no cuBLAS in the capture), so
        # the strict mode holds for this thread; other threads (data-pipeline workers, pinned-memory threads) may keep
        # calling into the runtime while the capture is open
        n0 = _lib.launch_count()
        # the kernel-layout weights are re-derived from the parameters INSIDE the forward graph (every replay sees the
        # parameters of that moment).  Those copies live in the graph's private pool and are only valid between the
        # forward and the backward replay of one step, so the engine-level cache (what eager forwards and the sampler
        # graphs use) is saved here and restored after the capture instead of being left pointing into the pool.
        keep = (self._wcache, self._wversion)
        try:
            return self._capture_train_graphs(xs, ts, cs, gs, x, sig, pool, g_f, g_b, n0, capture_mode)
        finally:
            # also after a FAILED capture: tensors created inside it were never computed, the cache must not point at them
            self._wcache, self._wversion = keep
            self._wgen += 1

    def _capture_train_graphs(self, xs, ts, cs, gs, x, sig, pool, g_f, g_b, n0, capture_mode):
        from .backward import BackwardProgram

        with _gc_paused(), torch.cuda.graph(g_f, pool=pool, capture_error_mode=capture_mode):
            self._wcache = None
            eps = self.forward(xs, ts, cs, train=True)
        n_l1 = _lib.launch_count()
        from .backward import grad_phase

        # flat gradient buffer: phase-1 parameters (decoder, up path, centre) first, then phase 2 - each phase's gradients are
        # packed at the end of ITS graph, so a data-parallel exchange of the first part overlaps the second graph
        by_name = dict(self.model.named_parameters())
        order = [n for n in by_name if grad_phase(n) == 1] + [n for n in by_name if grad_phase(n) == 2]
        sizes = [by_name[n].numel() for n in order]
        n1 = sum(by_name[n].numel() for n in order if grad_phase(n) == 1)
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=x.device)
        views = {n: v.view(by_name[n].shape) for v, n in zip(flat.split(sizes), order)}
        g_b2 = torch.cuda.CUDAGraph()
        bp = BackwardProgram(self)
        with _gc_paused(), torch.cuda.graph(g_b, pool=pool, capture_error_mode=capture_mode):
            bp.phase1(gs)
            first = [n for n in order if grad_phase(n) == 1]
            missing = [n for n in first if n not in bp.grads]
            if missing:
                raise RuntimeError(f"turbdiff_b200: phase 1 of the backward program left no gradient for {missing[:3]}")
            # all gradients of this phase packed into the flat buffer: the autograd glue then hands them out with a single
            # copy instead of one per tensor
            torch._foreach_copy_([views[n] for n in first], [bp.grads[n].reshape(by_name[n].shape) for n in first])
        with _gc_paused(), torch.cuda.graph(g_b2, pool=pool, capture_error_mode=capture_mode):
            grads, g_c = bp.phase2()
            second = [n for n in order if grad_phase(n) == 2]
            torch._foreach_copy_([views[n] for n in second], [grads[n].reshape(by_name[n].shape) for n in second])
        n2 = _lib.launch_count()
        pool_w = self._wcache  # kept alive by the returned record: the graphs hold raw addresses into it
        return {"sig": sig, "n_fwd": n_l1 - n0, "n_bwd": n2 - n_l1, "flat": flat, "order": order, "sizes": sizes, "n1": n1,
                "shapes": [by_name[n].shape for n in order], "x": xs, "t": ts, "c": cs, "g_eps": gs, "eps": eps, "fwd": g_f, "bwd": g_b,
                "bwd2": g_b2, "grads": grads, "g_c_local": g_c, "pool_w": pool_w}

    @staticmethod
    def to_ncdhw(v: View) -> torch.Tensor:
        """Interior of a halo-grid view as an fp32 NCDHW tensor (debug / tests)."""
        t = v.t[:, 1:-1, 1:-1, 1:-1, v.c0 : v.c0 + v.C]
        return t.permute(0, 4, 1, 2, 3).float().contiguous()
