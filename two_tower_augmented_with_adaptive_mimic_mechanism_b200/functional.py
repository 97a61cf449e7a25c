"""Tensor-level wrappers over the C ABI (include/ttam.h).  torch is used for device memory and
streams only; every op below is one or two launches of our own sm_100a kernels."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT, OPT, PREC, TensorList, check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, dtype, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a tensor")
    if not t.is_cuda:
        raise _lib.TtamError(f"{name} is on {t.device}; this package has no CPU path (CUDA tensors only)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")


def _rows2d(t: torch.Tensor, name: str):
    """2-D tensor with unit inner stride -> (ptr, ld)."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError(f"{name} must be 2-D with contiguous rows")
    return t.data_ptr(), (t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1]))


def _ptr(t):
    return None if t is None else t.data_ptr()


_ws_cache: dict = {}
_alloc_generation = [0]


def alloc_generation() -> int:
    """Bumped whenever a grow-only buffer that kernels address by raw pointer (scratch workspace, engine / tower
    buffers) is reallocated.  A captured CUDA graph holds the OLD pointers: holders of graphs compare the generation
    they captured at with this one before every replay and re-capture when it moved."""
    return _alloc_generation[0]


def note_realloc() -> None:
    _alloc_generation[0] += 1


_ws_scope = [""]


class ws_scope:
    """Scratch buffers requested inside this context are private to `name`: two branches of a step that run
    concurrently on different CUDA streams (user tower / item tower) must not share scratch memory."""

    def __init__(self, name: str) -> None:
        self.name = name

    def __enter__(self):
        self.prev = _ws_scope[0]
        _ws_scope[0] = self.name
        return self

    def __exit__(self, *a):
        _ws_scope[0] = self.prev


def workspace(nbytes: int, device, tag: str = "default") -> torch.Tensor:
    """Grow-only scratch buffer per (device, tag, scope) so that hot loops never allocate."""
    key = (str(device), tag, _ws_scope[0])
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        grow = int(nbytes) if buf is None else int(nbytes * 1.25)   # regrowth gets headroom (row counts drift per step)
        if buf is not None:
            note_realloc()
        buf = torch.empty(max(grow, 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# ---------------------------------------------------------------------------------------------
def gather_rows(table: torch.Tensor, idx: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    _chk(table, torch.float32, "table")
    _chk(idx, torch.int64, "idx")
    idx = idx.contiguous().view(-1)
    tp, ldt = _rows2d(table, "table")
    R, ncols = idx.numel(), table.shape[1]
    if out is None:
        out = torch.empty((R, ncols), dtype=torch.float32, device=table.device)
    op, ldo = _rows2d(out, "out")
    check(lib().ttam_gather_rows_f32(tp, ldt, table.shape[0], idx.data_ptr(), op, ldo, R, ncols, _stream()), "gather_rows")
    return out


def cast_bf16(src: torch.Tensor) -> torch.Tensor:
    _chk(src, torch.float32, "src")
    sp, lds = _rows2d(src, "src")
    dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    check(lib().ttam_cast_f32_to_bf16(sp, lds, dst.data_ptr(), dst.stride(0), src.shape[0], src.shape[1], _stream()), "cast_bf16")
    return dst


def round_tf32_(t: torch.Tensor) -> torch.Tensor:
    """In place: every element of the 2-D tensor (unit inner stride) becomes its nearest TF32-representable value."""
    _chk(t, torch.float32, "t")
    ptr, ld = _rows2d(t, "t")
    check(lib().ttam_round_tf32(ptr, ld, t.shape[0], t.shape[1], _stream()), "round_tf32")
    return t


def pad_cols(t: torch.Tensor, mult: int = 4, *, out=None, always_copy: bool = False) -> torch.Tensor:
    """View [R, C] of a zero-padded copy whose rows are a multiple of `mult` floats (16-byte aligned rows for the
    tensor-core loaders).  Returns `t` itself when it already qualifies."""
    R, Ccols = t.shape
    if not always_copy and Ccols % mult == 0 and t.stride(1) == 1 and t.stride(0) % mult == 0 and t.data_ptr() % 16 == 0:
        return t
    ld = (Ccols + mult - 1) // mult * mult
    if out is None or out.shape != (R, ld) or out.device != t.device:
        out = torch.zeros((R, ld), dtype=t.dtype, device=t.device)
    out[:, :Ccols].copy_(t)
    return out[:, :Ccols]


def _prec(precision: str, x_rounded: bool = False, w_rounded: bool = False, out_rounded: bool = False, wt: bool = False) -> int:
    return (PREC[precision] | (_lib.PREC_X_ROUNDED if x_rounded else 0) | (_lib.PREC_W_ROUNDED if w_rounded else 0)
            | (_lib.PREC_OUT_ROUNDED if out_rounded else 0) | (_lib.PREC_WT if wt else 0))


def prepare_weights(weights, want_rounded=True, want_transposed=True):
    """[(W [rows, cols])...] (at most 4) -> ([TF32-rounded copies], [TF32-rounded transposed copies]) in one launch."""
    import ctypes as C
    n = len(weights)
    dst = [torch.empty_like(w) if want_rounded else None for w in weights]
    dst_t = [torch.empty((w.shape[1], w.shape[0]), dtype=torch.float32, device=w.device) if want_transposed else None for w in weights]
    arr_p = lambda ts: (C.c_void_p * n)(*[None if t is None else t.data_ptr() for t in ts])
    arr_i = lambda vs: (C.c_int64 * n)(*vs)
    for w in weights:
        _chk(w, torch.float32, "weight")
    check(lib().ttam_prepare_weights(arr_p(weights), arr_i([w.stride(0) for w in weights]), arr_i([w.shape[0] for w in weights]),
                                     arr_i([w.shape[1] for w in weights]), arr_p(dst), arr_p(dst_t), n, _stream()), "prepare_weights")
    return dst, dst_t


def linear_fwd(x, w, bias=None, *, gather=None, act="none", out=None, dropout_p=0.0, seed=0, offset=0,
               state=None, precision="fp32", x_rounded=False, w_rounded=False, out_rounded=False):
    """x_rounded / w_rounded (tensor-core path): the operand was passed through round_tf32_ when it was laid out, the
    GEMM skips its rounding pass for it."""
    _chk(x, torch.float32, "x"); _chk(w, torch.float32, "w")
    xp, ldx = _rows2d(x, "x")
    wp, ldw = _rows2d(w, "w")          # [N,K], rows ldw apart (a padded view is fine)
    N, K = w.shape
    if x.shape[1] != K:
        raise ValueError(f"x has {x.shape[1]} columns, weight expects {K}")
    if gather is not None:
        _chk(gather, torch.int64, "gather")
        M = gather.numel()
    else:
        M = x.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    yp, ldy = _rows2d(out, "out")
    check(lib().ttam_linear_fwd(xp, ldx, _ptr(gather), wp, ldw, _ptr(bias), yp, ldy, M, N, K, ACT[act],
                                float(dropout_p), int(seed), int(offset), _ptr(state), _prec(precision, x_rounded, w_rounded, out_rounded),
                                _stream()), "linear_fwd")
    return out


def linear_dgrad(dy, w, *, out=None, aux=None, relu_mask=False, scale=1.0, accumulate=False, precision="fp32",
                 w_transposed=False, x_rounded=False, out_rounded=False):
    """w_transposed: `w` is the TF32-rounded transposed weight [K, N] from prepare_weights (TMA-fed kernel)."""
    _chk(dy, torch.float32, "dy"); _chk(w, torch.float32, "w")
    dyp, lddy = _rows2d(dy, "dy")
    N, K = (w.shape[1], w.shape[0]) if w_transposed else w.shape
    if w_transposed and not w.is_contiguous():
        raise ValueError("the transposed weight must be contiguous")
    M = dy.shape[0]
    if out is None:
        if accumulate:
            raise ValueError("accumulate needs an output buffer")
        out = torch.empty((M, K), dtype=torch.float32, device=dy.device)
    dxp, lddx = _rows2d(out, "out")
    auxp, ldaux = (None, 0)
    if relu_mask:
        auxp, ldaux = _rows2d(aux, "aux")
    check(lib().ttam_linear_dgrad(dyp, lddy, w.data_ptr(), dxp, lddx, auxp, ldaux, 1 if relu_mask else 0, float(scale),
                                  1 if accumulate else 0, M, N, K, _prec(precision, x_rounded, False, out_rounded, w_transposed),
                                  _stream()), "linear_dgrad")
    return out


def linear_wgrad(dy, x, *, gather=None, dw=None, db=None, accumulate=False, precision="fp32", want_bias=True,
                 x_rounded=False):
    _chk(dy, torch.float32, "dy"); _chk(x, torch.float32, "x")
    dyp, lddy = _rows2d(dy, "dy")
    xp, ldx = _rows2d(x, "x")
    M, N = dy.shape
    K = x.shape[1]
    if dw is None:
        dw = torch.empty((N, K), dtype=torch.float32, device=dy.device)
        accumulate = False
    if db is None and want_bias:
        db = torch.empty((N,), dtype=torch.float32, device=dy.device)
    nbytes = lib().ttam_linear_wgrad_workspace_bytes(M, N, K)
    ws = workspace(nbytes, dy.device, "wgrad")
    check(lib().ttam_linear_wgrad(dyp, lddy, xp, ldx, _ptr(gather), dw.data_ptr(), _ptr(db), M, N, K,
                                  1 if accumulate else 0, ws.data_ptr(), ws.numel(), _prec(precision, x_rounded), _stream()),
          "linear_wgrad")
    return dw, db


# ---------------------------------------------------------------------------------------------
# bag form of a feature matrix (csrc/bag.cu)
# ---------------------------------------------------------------------------------------------
BAG_MAX_NNZ = 64      # sparse entries per row the kernels stage (csrc/bag.cu kMaxNnz)
BAG_MAX_TAIL = 8      # dense trailing columns (kMaxTail)
BAG_WGRAD_MAX_MEAN_NNZ = 8.0


class BagMatrix:
    """CSR + dense-tail layout of a feature matrix X [N, F] (reference features.py:242-252: one-/multi-hot category and
    author columns, then a few z-scored numerics that are non-zero in every row).

      tail_start : the trailing columns [tail_start, F) whose density is >= 1/2 (at most 8) are kept as a dense [N, T] block
      rowptr     : int64 [N + 1], entries of row r are entries[rowptr[r] : rowptr[r + 1]] in column order
      entries    : int32 [nnz, 2] = (column, fp32 value bits): one 8-byte load per non-zero
    Built with torch index ops, once per feature matrix (layout preparation, like the padded copy of the dense path).
    `BagMatrix.build` returns None when the matrix is too dense for the bag kernels (a row with more than 64 non-zeros
    outside the tail): the caller keeps the dense GEMM path."""

    def __init__(self, rowptr, entries, tail, tail_start, shape, mean_nnz, max_nnz):
        self.rowptr, self.entries, self.tail, self.tail_start, self.shape, self.mean_nnz = rowptr, entries, tail, tail_start, shape, mean_nnz
        self.max_nnz = int(max_nnz)          # most CSR entries in one row: sizes the kernels' tile ring
        # The forward beats the dense GEMM at any density the layout admits; the weight-gradient scatter costs ~20
        # instructions per (entry, 64 hidden columns) and loses to the dense tensor-core GEMM beyond ~8 sparse entries a row
        # (measured: 49 152 rows x 3.5 entries: 69 us against 107; 8 192 rows x 26 entries: 64 us against 40).
        self.wgrad = (mean_nnz - (shape[1] - tail_start)) <= BAG_WGRAD_MAX_MEAN_NNZ
        self.T = shape[1] - tail_start

    @staticmethod
    def build(X: torch.Tensor, chunk_rows: int = 1 << 18):
        if X.dtype != torch.float32 or X.dim() != 2:
            raise ValueError("BagMatrix.build needs a 2-D float32 matrix")
        N, Fd = X.shape
        if N == 0 or Fd == 0 or N >= 2 ** 31:      # the kernels keep row ids as int32 in shared memory
            return None
        # dense tail: the longest suffix of columns (<= 8) that are non-zero in at least half of the rows
        probe = X[: min(N, 1 << 16), max(0, Fd - BAG_MAX_TAIL):]
        dens = (probe != 0).float().mean(0).cpu().tolist()
        T = 0
        for d in reversed(dens):
            if d >= 0.5:
                T += 1
            else:
                break
        tail_start = Fd - T
        counts, cols, vals = [], [], []
        max_nnz = 0
        for s in range(0, N, chunk_rows):
            blk = X[s:s + chunk_rows, :tail_start]
            nz = blk != 0
            cnt = nz.sum(1)
            max_nnz = max(max_nnz, int(cnt.max()) if cnt.numel() else 0)
            if max_nnz > BAG_MAX_NNZ:
                return None
            rc = nz.nonzero()                      # row-major order: columns ascending inside a row
            counts.append(cnt)
            cols.append(rc[:, 1].to(torch.int32))
            vals.append(blk[nz])
        cnt = torch.cat(counts)
        rowptr = torch.zeros(N + 1, dtype=torch.int64, device=X.device)
        torch.cumsum(cnt, 0, out=rowptr[1:])
        col = torch.cat(cols)
        val = torch.cat(vals)
        if col.numel() >= 2 ** 32:                  # ... and entry offsets as uint32
            return None
        entries = torch.empty((max(col.numel(), 1), 2), dtype=torch.int32, device=X.device)
        if col.numel():
            entries[:, 0] = col
            entries[:, 1] = val.view(torch.int32)
        tail = X[:, tail_start:].contiguous() if T > 0 else None
        return BagMatrix(rowptr, entries, tail, tail_start, (N, Fd), float(col.numel()) / N + T, max_nnz)


def bag_supported(H: int, Fdim: int, T: int = 0, max_nnz: int = BAG_MAX_NNZ) -> bool:
    return bool(lib().ttam_bag_supported(int(H), int(Fdim), int(T), int(max_nnz)))


def bag_linear_fwd(bag: BagMatrix, gather, w, bias=None, *, act="none", out=None, dropout_p=0.0, seed=0, offset=0, state=None,
                   round_tf32_out=False):
    """out[r] = act(bias + sum_j X[gather[r], j] w[:, j]) from the bag form of X (fp32 FMA)."""
    _chk(w, torch.float32, "w")
    wp, ldw = _rows2d(w, "w")
    H, Fd = w.shape
    if Fd != bag.shape[1]:
        raise ValueError(f"weight expects {Fd} features, the bag matrix has {bag.shape[1]}")
    if gather is not None:
        _chk(gather, torch.int64, "gather")
        R = gather.numel()
    else:
        R = bag.shape[0]
    if out is None:
        out = torch.empty((R, H), dtype=torch.float32, device=w.device)
    yp, ldy = _rows2d(out, "out")
    L = lib()
    ws = workspace(L.ttam_bag_linear_workspace_bytes(R, H, Fd), w.device, "bag")
    check(L.ttam_bag_linear_fwd(bag.rowptr.data_ptr(), bag.entries.data_ptr(), _ptr(bag.tail), bag.T, bag.tail_start, bag.max_nnz, _ptr(gather), R,
                                wp, ldw, _ptr(bias), yp, ldy, H, Fd, ACT[act], float(dropout_p), int(seed), int(offset), _ptr(state),
                                1 if round_tf32_out else 0, ws.data_ptr(), ws.numel(), _stream()), "bag_linear_fwd")
    return out


def bag_linear_wgrad(bag: BagMatrix, gather, dh, *, dw=None, db=None, accumulate=False, want_bias=True, precision="fp32"):
    """dw[h, j] (+)= sum_r dh[r, h] X[gather[r], j];  db[h] (+)= sum_r dh[r, h].  Deterministic.
    precision="tf32": the tensor-core kernel (CSR rows expanded into the MMA operand tile; TF32 products, exact fp32 bias
    gradient) where the shape is covered, else - and for "fp32" - the fp32 scatter kernel."""
    _chk(dh, torch.float32, "dh")
    dp, lddh = _rows2d(dh, "dh")
    R, H = dh.shape
    Fd = bag.shape[1]
    if dw is None:
        dw = torch.empty((H, Fd), dtype=torch.float32, device=dh.device)
        accumulate = False
    if db is None and want_bias:
        db = torch.empty((H,), dtype=torch.float32, device=dh.device)
    L = lib()
    if precision == "tf32" and R > 0:
        ws = workspace(L.ttam_bag_linear_wgrad_tc_workspace_bytes(R, H, Fd), dh.device, "bag")
        rc = L.ttam_bag_linear_wgrad_tc(bag.rowptr.data_ptr(), bag.entries.data_ptr(), _ptr(bag.tail), bag.T, bag.tail_start, _ptr(gather), R,
                                        dp, lddh, dw.data_ptr(), dw.stride(0), _ptr(db), H, Fd, 1 if accumulate else 0, ws.data_ptr(),
                                        ws.numel(), _stream())
        if rc == 0:
            return dw, db
        if rc != 1:
            check(rc, "bag_linear_wgrad_tc")
    ws = workspace(L.ttam_bag_linear_workspace_bytes(R, H, Fd), dh.device, "bag")
    check(L.ttam_bag_linear_wgrad(bag.rowptr.data_ptr(), bag.entries.data_ptr(), _ptr(bag.tail), bag.T, bag.tail_start, bag.max_nnz, _ptr(gather), R,
                                  dp, lddh, dw.data_ptr(), dw.stride(0), _ptr(db), H, Fd, 1 if accumulate else 0, ws.data_ptr(),
                                  ws.numel(), _stream()), "bag_linear_wgrad")
    return dw, db


def new_step_state(device, step: int = 0, rng_offset: int = 0) -> torch.Tensor:
    """Device-resident ttam_step_state {int32 step; int32 pad; uint64 rng_offset} as an int64[2] tensor."""
    return torch.tensor([int(step) & 0xFFFFFFFF, int(rng_offset)], dtype=torch.int64, device=device)


def advance_step(state: torch.Tensor, rng_stride: int = 1 << 32) -> None:
    check(lib().ttam_advance_step(state.data_ptr(), int(rng_stride), _stream()), "advance_step")


def act_fwd(pre, *, act, out=None, dropout_p=0.0, seed=0, offset=0, state=None):
    _chk(pre, torch.float32, "pre")
    if out is None:
        out = torch.empty_like(pre)
    check(lib().ttam_act_fwd(pre.data_ptr(), out.data_ptr(), pre.numel(), pre.shape[-1], ACT[act], float(dropout_p),
                             int(seed), int(offset), _ptr(state), _stream()), "act_fwd")
    return out


def act_bwd(dy, pre, *, act, out=None, dropout_p=0.0, seed=0, offset=0, state=None):
    _chk(dy, torch.float32, "dy")
    if out is None:
        out = torch.empty_like(pre)
    check(lib().ttam_act_bwd(dy.data_ptr(), pre.data_ptr(), out.data_ptr(), pre.numel(), pre.shape[-1], ACT[act],
                             float(dropout_p), int(seed), int(offset), _ptr(state), _stream()), "act_bwd")
    return out


def gate_fwd(z, pre2, *, g, t=None, o=None, q=None, aug_table=None, idx=None):
    """g = sigmoid(pre2); t = g*e + (1-g)*f with [e;f] = z; o = t + aug[idx]; q = aug[idx].  Outputs are caller-provided."""
    _chk(z, torch.float32, "z"); _chk(pre2, torch.float32, "pre2")
    R, D = pre2.shape
    for name, ten in (("z", z), ("pre2", pre2), ("g", g), ("t", t), ("o", o), ("q", q)):
        if ten is not None and not ten.is_contiguous():
            raise ValueError(f"gate_fwd: {name} must be contiguous")
    check(lib().ttam_gate_fwd(z.data_ptr(), pre2.data_ptr(), _ptr(aug_table), 0 if aug_table is None else aug_table.shape[0],
                              _ptr(idx), g.data_ptr(), _ptr(t), _ptr(o), _ptr(q), R, D, _stream()), "gate_fwd")
    return g, t, o, q


def gate_bwd(dt, z, g, *, dpre2, dz):
    R, D = g.shape
    for name, ten in (("dt", dt), ("z", z), ("g", g), ("dpre2", dpre2), ("dz", dz)):
        if not ten.is_contiguous():
            raise ValueError(f"gate_bwd: {name} must be contiguous")
    check(lib().ttam_gate_bwd(dt.data_ptr(), z.data_ptr(), g.data_ptr(), dpre2.data_ptr(), dz.data_ptr(), R, D, _stream()), "gate_bwd")
    return dpre2, dz


def augment_fwd(t, aug_table, idx, *, out, q_out=None):
    """out = t + aug_table[idx]; q_out = aug_table[idx]."""
    _chk(t, torch.float32, "t"); _chk(aug_table, torch.float32, "aug_table"); _chk(idx, torch.int64, "idx")
    R, D = t.shape
    for name, ten in (("t", t), ("out", out), ("q_out", q_out), ("idx", idx)):
        if ten is not None and not ten.is_contiguous():
            raise ValueError(f"augment_fwd: {name} must be contiguous")
    if aug_table.shape[1] != D or not aug_table.is_contiguous():
        raise ValueError("augment_fwd: table must be contiguous [N, D]")
    check(lib().ttam_augment_fwd(t.data_ptr(), aug_table.data_ptr(), aug_table.shape[0], idx.data_ptr(),
                                 out.data_ptr(), _ptr(q_out), R, D, _stream()), "augment_fwd")
    return out, q_out


def loss_fwd_bwd(o_u, o_i, *, t_u=None, t_p=None, q_u=None, q_p=None, lambda_u=0.0, lambda_i=0.0, backward=True,
                 out=None, batch_fraction=1.0):
    """o_i = [positives (B rows); negatives (B*N rows, [B,N] row-major)].  Returns (loss[4], do_u, do_i, dq_u, dq_p)."""
    _chk(o_u, torch.float32, "o_u"); _chk(o_i, torch.float32, "o_i")
    B, D = o_u.shape
    N = o_i.shape[0] // B - 1
    dev = o_u.device
    mimic = q_u is not None
    if out is not None:      # (loss[4], do_u, do_i, dq_u, dq_p) provided by the caller
        loss, do_u, do_i, dq_u, dq_p = out
    else:
        loss = torch.empty(4, dtype=torch.float32, device=dev)
        do_u = torch.empty_like(o_u) if backward else None
        do_i = torch.empty_like(o_i) if backward else None
        dq_u = torch.empty_like(o_u) if (backward and mimic) else None
        dq_p = torch.empty_like(o_u) if (backward and mimic) else None
    ws = workspace(lib().ttam_loss_workspace_bytes(B), dev, "loss")
    check(lib().ttam_loss_fwd_bwd(o_u.data_ptr(), o_i.data_ptr(), _ptr(t_u), _ptr(t_p), _ptr(q_u), _ptr(q_p),
                                  float(lambda_u), float(lambda_i), loss.data_ptr(), _ptr(do_u), _ptr(do_i), _ptr(dq_u),
                                  _ptr(dq_p), B, N, D, float(batch_fraction), ws.data_ptr(), ws.numel(), _stream()), "loss_fwd_bwd")
    return loss, do_u, do_i, dq_u, dq_p


def loss_aug_supported(N: int, D: int) -> bool:
    return bool(lib().ttam_loss_aug_supported(int(N), int(D)))


def loss_aug_fwd_bwd(t_u, t_i, aug_u, aug_i, users, items, *, mimic=True, lambda_u=0.0, lambda_i=0.0, backward=True, out=None,
                     batch_fraction=1.0):
    """Augmentation add + loss in one pass: o = t + aug[idx] never leaves the registers (csrc/rows.cu loss_aug_vec_kernel).
    t_i / items = [positives (B rows); negatives (B*N rows, [B,N] row-major)].  Returns (loss[4], do_u, do_i, dq_u, dq_p)."""
    for t, n in ((t_u, "t_u"), (t_i, "t_i"), (aug_u, "aug_u"), (aug_i, "aug_i")):
        _chk(t, torch.float32, n)
    _chk(users, torch.int64, "users"); _chk(items, torch.int64, "items")
    B, D = t_u.shape
    N = t_i.shape[0] // B - 1
    if not (t_u.is_contiguous() and t_i.is_contiguous() and aug_u.is_contiguous() and aug_i.is_contiguous()):
        raise ValueError("loss_aug_fwd_bwd needs contiguous rows")
    dev = t_u.device
    if out is not None:
        loss, do_u, do_i, dq_u, dq_p = out
    else:
        loss = torch.empty(4, dtype=torch.float32, device=dev)
        do_u = torch.empty_like(t_u) if backward else None
        do_i = torch.empty_like(t_i) if backward else None
        dq_u = torch.empty_like(t_u) if (backward and mimic) else None
        dq_p = torch.empty_like(t_u) if (backward and mimic) else None
    ws = workspace(lib().ttam_loss_workspace_bytes(B), dev, "loss")
    check(lib().ttam_loss_aug_fwd_bwd(t_u.data_ptr(), t_i.data_ptr(), aug_u.data_ptr(), aug_u.shape[0], aug_i.data_ptr(), aug_i.shape[0],
                                      users.data_ptr(), items.data_ptr(), 1 if mimic else 0, float(lambda_u), float(lambda_i),
                                      loss.data_ptr(), _ptr(do_u), _ptr(do_i), _ptr(dq_u), _ptr(dq_p), B, N, D, float(batch_fraction),
                                      ws.data_ptr(), ws.numel(), _stream()), "loss_aug_fwd_bwd")
    return loss, do_u, do_i, dq_u, dq_p


def loss_slots_fwd_bwd(t_u: list, q_u: list, t_i: list, q_i: list, a_u: list, b_u: list, a_i: list, b_i: list, cap_u: int, cap_i: int,
                       slot_of_u, slot_of_i, B: int, N: int, D: int, *, lambda_u=0.0, lambda_i=0.0, loss=None, batch_fraction=1.0):
    """The sharded step's loss with the row exchange folded in (csrc/rows.cu loss_slots_vec_kernel): the eight lists hold one
    device address per rank (the owners' row buffers and receive buffers, local or NVLink peer mappings)."""
    _chk(slot_of_u, torch.int64, "slot_of_u"); _chk(slot_of_i, torch.int64, "slot_of_i")
    dev = slot_of_u.device
    if loss is None:
        loss = torch.empty(4, dtype=torch.float32, device=dev)
    W = len(t_u)
    ws = workspace(lib().ttam_loss_workspace_bytes(B), dev, "loss")
    arrs = [_ptr_array(x) for x in (t_u, q_u, t_i, q_i, a_u, b_u, a_i, b_i)]
    check(lib().ttam_loss_slots_fwd_bwd(*arrs, W, int(cap_u), int(cap_i), slot_of_u.data_ptr(), slot_of_i.data_ptr(), float(lambda_u),
                                        float(lambda_i), loss.data_ptr(), int(B), int(N), int(D), float(batch_fraction), ws.data_ptr(),
                                        ws.numel(), _stream()), "loss_slots_fwd_bwd")
    return loss


def inbatch_loss_fwd_bwd(o_u, o_p, *, t_u=None, t_p=None, q_u=None, q_p=None, lambda_u=0.0, lambda_i=0.0, backward=True, out=None,
                         precision="fp32"):
    """In-batch softmax loss (extension; see ttam.h): o_u, o_p [B, D].  Returns (loss[4], do_u, do_p, dq_u, dq_p)."""
    _chk(o_u, torch.float32, "o_u"); _chk(o_p, torch.float32, "o_p")
    B, D = o_u.shape
    dev = o_u.device
    mimic = q_u is not None
    for name, ten in (("o_u", o_u), ("o_p", o_p), ("t_u", t_u), ("t_p", t_p), ("q_u", q_u), ("q_p", q_p)):
        if ten is not None and not ten.is_contiguous():
            raise ValueError(f"inbatch_loss_fwd_bwd: {name} must be contiguous")
    if out is not None:
        loss, do_u, do_p, dq_u, dq_p = out
    else:
        loss = torch.empty(4, dtype=torch.float32, device=dev)
        do_u = torch.empty_like(o_u) if backward else None
        do_p = torch.empty_like(o_p) if backward else None
        dq_u = torch.empty_like(o_u) if (backward and mimic) else None
        dq_p = torch.empty_like(o_u) if (backward and mimic) else None
    L = lib()
    ws = workspace(L.ttam_inbatch_loss_workspace_bytes(B, D), dev, "inbatch")
    check(L.ttam_inbatch_loss_fwd_bwd(o_u.data_ptr(), o_p.data_ptr(), _ptr(t_u), _ptr(t_p), _ptr(q_u), _ptr(q_p), float(lambda_u),
                                      float(lambda_i), loss.data_ptr(), _ptr(do_u), _ptr(do_p), _ptr(dq_u), _ptr(dq_p), B, D,
                                      PREC[precision], ws.data_ptr(), ws.numel(), _stream()), "inbatch_loss_fwd_bwd")
    return loss, do_u, do_p, dq_u, dq_p


def category_alignment(item_idx, emb, cat_tensor, n_categories: int, major: int, *, lambda_c: float = 1.0,
                       loss_out=None, grad_a=None, grad_b=None, B: int = 0):
    """Category-alignment loss of `emb` [R, D] (rows = item_idx) and its gradient (see ttam.h).  Returns the raw loss
    as a 1-element device tensor; lambda_c * loss is ADDED to loss_out[0], lambda_c * gradient to grad_a (and grad_b)."""
    _chk(emb, torch.float32, "emb"); _chk(item_idx, torch.int64, "item_idx"); _chk(cat_tensor, torch.int64, "cat_tensor")
    R, D = emb.shape
    if grad_a is None:
        grad_a = torch.zeros_like(emb)
    cal = torch.zeros(1, dtype=torch.float32, device=emb.device)
    L = lib()
    ws = workspace(L.ttam_category_alignment_workspace_bytes(R, D, int(n_categories)), emb.device, "catalign")
    check(L.ttam_category_alignment(item_idx.data_ptr(), R, emb.data_ptr(), D, cat_tensor.data_ptr(), cat_tensor.numel(),
                                    int(n_categories), int(major), float(lambda_c), _ptr(loss_out), cal.data_ptr(),
                                    grad_a.data_ptr(), _ptr(grad_b), int(B), ws.data_ptr(), ws.numel(), _stream()),
          "category_alignment")
    return cal, grad_a


# ---------------------------------------------------------------------------------------------
def sort_rows(idx: torch.Tensor, num_rows: int, *, sorted_idx=None, perm=None):
    """Stable sort of the touched row ids -> (sorted ids, original positions int32)."""
    _chk(idx, torch.int64, "idx")
    idx = idx.contiguous().view(-1)
    R = idx.numel()
    if sorted_idx is None:
        sorted_idx = torch.empty_like(idx)
    if perm is None:
        perm = torch.empty(R, dtype=torch.int32, device=idx.device)
    ws = workspace(lib().ttam_sort_workspace_bytes(R), idx.device, "sort")
    check(lib().ttam_sort_rows(idx.data_ptr(), R, num_rows, sorted_idx.data_ptr(), perm.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "sort_rows")
    return sorted_idx, perm


def unique_rows(sorted_idx: torch.Tensor) -> torch.Tensor:
    R = sorted_idx.numel()
    out = torch.empty_like(sorted_idx)
    n = torch.zeros(1, dtype=torch.int64, device=sorted_idx.device)
    ws = workspace(lib().ttam_sort_workspace_bytes(R), sorted_idx.device, "sort")
    check(lib().ttam_unique_rows(sorted_idx.data_ptr(), R, out.data_ptr(), n.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "unique_rows")
    return out[: int(n.item())]


def find_long_segments(sorted_idx: torch.Tensor, *, out=None) -> torch.Tensor:
    """int32 list {count, head positions...} of the segments of `sorted_idx` longer than 16 rows (see ttam.h)."""
    R = sorted_idx.numel()
    n = lib().ttam_long_segments_bytes(R) // 4
    if out is None or out.numel() < n:
        out = torch.empty(n, dtype=torch.int32, device=sorted_idx.device)
    check(lib().ttam_find_long_segments(sorted_idx.data_ptr(), R, out.data_ptr(), _stream()), "find_long_segments")
    return out


def _grad_src(grad_a, grad_b):
    ap, lda = _rows2d(grad_a, "grad_a")
    n_a = grad_a.shape[0]
    if grad_b is None:
        return ap, lda, n_a, None, 0
    bp, ldb = _rows2d(grad_b, "grad_b")
    return ap, lda, n_a, bp, ldb


def sparse_adam_rows(p, m, v, sorted_idx, perm, grad_a, grad_b=None, *, lr, betas=(0.9, 0.999), eps=1e-8, step=1,
                     scalars=None, state=None, long_list=None):
    ap, lda, n_a, bp, ldb = _grad_src(grad_a, grad_b)
    R = sorted_idx.numel()
    check(lib().ttam_sparse_adam_rows(p.data_ptr(), m.data_ptr(), v.data_ptr(), p.shape[1], sorted_idx.data_ptr(),
                                      perm.data_ptr(), R, ap, lda, n_a, bp, ldb, _ptr(scalars), float(lr), float(betas[0]),
                                      float(betas[1]), float(eps), int(step), _ptr(state), _ptr(long_list), _stream()),
          "sparse_adam_rows")


def lazy_rows(kind, p, m, v, last_step, sorted_idx, perm, grad_a, grad_b=None, *, scalars, lr, weight_decay=0.0,
              betas=(0.9, 0.999), eps=1e-8, momentum=0.0, step=1, state=None, long_list=None):
    ap, lda, n_a, bp, ldb = _grad_src(grad_a, grad_b)
    R = sorted_idx.numel()
    check(lib().ttam_lazy_rows(OPT[kind], p.data_ptr(), _ptr(m), _ptr(v), last_step.data_ptr(), p.shape[1],
                               sorted_idx.data_ptr(), perm.data_ptr(), R, ap, lda, n_a, bp, ldb, _ptr(scalars),
                               float(lr), float(weight_decay), float(betas[0]), float(betas[1]), float(eps),
                               float(momentum), int(step), _ptr(state), _ptr(long_list), _stream()), "lazy_rows")


def lazy_catchup(kind, p, m, v, last_step, sorted_idx, *, scalars, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8,
                 momentum=0.0, step=1, state=None):
    """Replay the zero-gradient steps of the rows in sorted_idx up to step-1 (call before the forward of `step`)."""
    check(lib().ttam_lazy_catchup(OPT[kind], p.data_ptr(), _ptr(m), _ptr(v), last_step.data_ptr(), p.shape[1],
                                  sorted_idx.data_ptr(), sorted_idx.numel(), _ptr(scalars), float(lr), float(weight_decay),
                                  float(betas[0]), float(betas[1]), float(eps), float(momentum), int(step), _ptr(state),
                                  _stream()), "lazy_catchup")


def lazy_flush(kind, p, m, v, last_step, *, scalars, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8,
               momentum=0.0, step=1, state=None):
    check(lib().ttam_lazy_flush(OPT[kind], p.data_ptr(), _ptr(m), _ptr(v), last_step.data_ptr(), p.shape[0], p.shape[1],
                                _ptr(scalars), float(lr), float(weight_decay), float(betas[0]), float(betas[1]),
                                float(eps), float(momentum), int(step), _ptr(state), _stream()), "lazy_flush")


def dense_step(kind, params, grads, ms, vs, *, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, momentum=0.0, step=1,
               scalars=None, state=None):
    """AdamW / Adam / SGD over lists of small contiguous fp32 tensors, MAX_TENSORS per launch."""
    n = len(params)
    for s in range(0, n, _lib.MAX_TENSORS):
        tl = TensorList()
        cnt = min(_lib.MAX_TENSORS, n - s)
        tl.count = cnt
        for j in range(cnt):
            p, g = params[s + j], grads[s + j]
            if not (p.is_contiguous() and g.is_contiguous()):
                raise ValueError("dense_step needs contiguous tensors")
            tl.p[j] = p.data_ptr(); tl.g[j] = g.data_ptr()
            tl.m[j] = _ptr(ms[s + j]) if ms is not None else None
            tl.v[j] = _ptr(vs[s + j]) if vs is not None else None
            tl.numel[j] = p.numel()
        check(lib().ttam_dense_step(OPT[kind], C.byref(tl), _ptr(scalars), float(lr), float(weight_decay), float(betas[0]),
                                    float(betas[1]), float(eps), float(momentum), int(step), _ptr(state), _stream()), "dense_step")


def adam_scalar_table(max_step: int, lr: float, betas=(0.9, 0.999), device="cuda") -> torch.Tensor:
    """scalars[4t] = lr/(1-b1^t), [4t+1] = sqrt(1-b2^t), [4t+2] = lr*sqrt(1-b2^t)/(1-b1^t) (SparseAdam),
    [4t+3] = b2f^(t/2)/sqrt(1-b2^t) with b2f = fp32(b2), the factor the reference's exp_avg_sq.mul_(beta2) multiplies by
    (zero-gradient replay of the lazily-updated tables, csrc/optim.cu replay()); computed in float64 like torch does
    (Python doubles), stored fp32."""
    t = torch.arange(0, max_step + 1, dtype=torch.float64)
    b1, b2 = betas
    bc1 = 1.0 - torch.pow(torch.tensor(b1, dtype=torch.float64), t)
    bc2 = 1.0 - torch.pow(torch.tensor(b2, dtype=torch.float64), t)
    bc1[0] = 1.0
    b2f = float(torch.tensor(b2, dtype=torch.float32))
    h = torch.pow(torch.tensor(b2f, dtype=torch.float64), t / 2.0) / torch.sqrt(bc2.clamp_min(1e-300))
    h[0] = 0.0
    tab = torch.stack([lr / bc1, torch.sqrt(bc2), lr * torch.sqrt(bc2) / bc1, h], dim=1).to(torch.float32).contiguous()
    return tab.view(-1).to(device)


# ---------------------------------------------------------------------------------------------
TOPK_F32_MAX_K = 1024     # csrc/topk.cu: kSelCap - kChunk
TOPK_BF16_MAX_K = 128     # csrc/topk_tc.cu


def score_pairs(q: torch.Tensor, items: torch.Tensor, cand: torch.Tensor) -> torch.Tensor:
    """out[r, c] = canonical fp32 score of q[r] against items[cand[r, c]] (-inf where cand < 0)."""
    _chk(q, torch.float32, "q"); _chk(items, torch.float32, "items"); _chk(cand, torch.int64, "cand")
    q, items, cand = q.contiguous(), items.contiguous(), cand.contiguous()
    R, C = cand.shape
    out = torch.empty((R, C), dtype=torch.float32, device=q.device)
    check(lib().ttam_score_pairs(q.data_ptr(), items.data_ptr(), cand.data_ptr(), R, C, q.shape[1], items.shape[0],
                                 out.data_ptr(), _stream()), "score_pairs")
    return out


def topk(q: torch.Tensor, items: torch.Tensor, k: int, *, id_offset: int = 0, after=None):
    """Exact inner-product top-k in canonical (-score,+id) order.  fp32 -> SIMT path; bf16 -> tcgen05 path.
    after = (scores [Q] fp32, ids [Q] int64): fp32 only - query r admits only items that sort strictly after
    (scores[r], ids[r]) (ids[r] < 0: all of them); the next page of a ranking."""
    if q.dtype != items.dtype:
        raise ValueError("q and items must have the same dtype")
    if not (q.is_cuda and items.is_cuda):
        raise _lib.TtamError("topk needs CUDA tensors; there is no CPU path")
    q = q.contiguous(); items = items.contiguous()
    Q, D = q.shape
    N = items.shape[0]
    k_eff = min(k, N)
    ids = torch.empty((Q, k_eff), dtype=torch.int64, device=q.device)
    scores = torch.empty((Q, k_eff), dtype=torch.float32, device=q.device)
    L = lib()
    if after is not None and q.dtype != torch.float32:
        raise ValueError("paged search (after=) is available on the fp32 path only")
    if q.dtype == torch.float32:
        ws = workspace(L.ttam_topk_f32_workspace_bytes(Q, N, D, k_eff), q.device, "topk")
        a_s, a_i = (None, None) if after is None else (after[0].contiguous(), after[1].contiguous())
        if after is not None:
            _chk(a_s, torch.float32, "after scores"); _chk(a_i, torch.int64, "after ids")
        check(L.ttam_topk_f32_after(q.data_ptr(), items.data_ptr(), Q, N, D, k_eff, id_offset, _ptr(a_s), _ptr(a_i),
                                    ids.data_ptr(), scores.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "topk_f32")
    elif q.dtype == torch.bfloat16:
        ws = workspace(L.ttam_topk_bf16_workspace_bytes(Q, N, D, k_eff), q.device, "topk")
        check(L.ttam_topk_bf16(q.data_ptr(), items.data_ptr(), Q, N, D, k_eff, id_offset, ids.data_ptr(), scores.data_ptr(),
                               ws.data_ptr(), ws.numel(), _stream()), "topk_bf16")
    else:
        raise ValueError(f"unsupported dtype {q.dtype}")
    return ids, scores


TOPK_TC_MAX_D = 256       # csrc/topk_tc.cu: kMaxD (both tensor-core paths)


def split_bf16x3(x: torch.Tensor, *, item_layout: bool) -> torch.Tensor:
    """fp32 [R, D] -> bf16 [R, 2*ceil16(D)] = [hi | lo], hi = bf16(x), lo = bf16(x - hi): the operand of `topk_f32_tc`, whose
    MMA schedule forms the three products qh.xh + ql.xh + qh.xl from the two copies (same layout for items and queries)."""
    _chk(x, torch.float32, "x")
    x = x.contiguous()
    R, D = x.shape
    out = torch.empty((R, int(lib().ttam_split_bf16x3_cols(D))), dtype=torch.bfloat16, device=x.device)
    check(lib().ttam_split_bf16x3(x.data_ptr(), R, D, 1 if item_layout else 0, out.data_ptr(), _stream()), "split_bf16x3")
    return out


def topk_f32_tc(q: torch.Tensor, items: torch.Tensor, items_split: torch.Tensor, k: int, *, id_offset: int = 0):
    """`topk` for fp32 operands with the candidate pass on the tensor cores (k <= 128, D <= 256): same ids and scores,
    bit for bit, as the SIMT path.  items_split = split_bf16x3(items, item_layout=True), kept by the index."""
    _chk(q, torch.float32, "q"); _chk(items, torch.float32, "items"); _chk(items_split, torch.bfloat16, "items_split")
    if not (q.is_cuda and items.is_cuda):
        raise _lib.TtamError("topk needs CUDA tensors; there is no CPU path")
    q = q.contiguous(); items = items.contiguous()
    Q, D = q.shape
    N = items.shape[0]
    k_eff = min(k, N)
    q_split = split_bf16x3(q, item_layout=False)
    ids = torch.empty((Q, k_eff), dtype=torch.int64, device=q.device)
    scores = torch.empty((Q, k_eff), dtype=torch.float32, device=q.device)
    L = lib()
    ws = workspace(L.ttam_topk_f32_tc_workspace_bytes(Q, N, D, k_eff), q.device, "topk")
    check(L.ttam_topk_f32_tc(q.data_ptr(), items.data_ptr(), q_split.data_ptr(), items_split.data_ptr(), Q, N, D, k_eff,
                             id_offset, ids.data_ptr(), scores.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "topk_f32_tc")
    return ids, scores


def topk_merge(ids: torch.Tensor, scores: torch.Tensor, k_out: int):
    """ids/scores [Q, parts, K_in] -> best k_out per query under (-score,+id)."""
    Q, parts, k_in = ids.shape
    out_i = torch.empty((Q, k_out), dtype=torch.int64, device=ids.device)
    out_s = torch.empty((Q, k_out), dtype=torch.float32, device=ids.device)
    check(lib().ttam_topk_merge(ids.contiguous().data_ptr(), scores.contiguous().data_ptr(), Q, parts, k_in, k_out,
                                out_i.data_ptr(), out_s.data_ptr(), _stream()), "topk_merge")
    return out_i, out_s


# ---- fixed-capacity slot route of the row-sharded step (csrc/slots.cu) -----------------------------------------------
def slot_plan(idx: torch.Tensor, world: int, cap: int, *, send_idx, slot_of, req_of, flag) -> None:
    _chk(idx, torch.int64, "idx")
    idx = idx.contiguous().view(-1)
    R = idx.numel()
    if send_idx.numel() < world * cap or slot_of.numel() < R or req_of.numel() < world * cap:
        raise ValueError("slot_plan: output buffers too small")
    ws = workspace(lib().ttam_slot_plan_workspace_bytes(R, world), idx.device, "slots")
    check(lib().ttam_slot_plan(idx.data_ptr(), R, world, cap, send_idx.data_ptr(), slot_of.data_ptr(), req_of.data_ptr(),
                               flag.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "slot_plan")


def _ptr_array(ptrs):
    import ctypes as C
    return None if ptrs is None else (C.c_void_p * len(ptrs))(*ptrs)


def slot_unpack(t_src: list, q_src, ld_src: int, cap: int, slot_of: torch.Tensor, D: int, *, t_out=None, q_out=None,
                o_out=None) -> None:
    """t_src / q_src: one device address per owner (ints).  Outputs [R, D] fp32, request order."""
    R = slot_of.numel()
    check(lib().ttam_slot_unpack(_ptr_array(t_src), _ptr_array(q_src), ld_src, len(t_src), cap, slot_of.data_ptr(), R, D,
                                 _ptr(t_out), _ptr(q_out), _ptr(o_out), _stream()), "slot_unpack")


def slot_pack(a: torch.Tensor, b0, b1, req_of: torch.Tensor, cap: int, a_dst: list, b_dst, ld_dst: int) -> None:
    """a [R, D]; b rows: b0[r] for r < b0.shape[0], b1[r] beyond (either may be None).  a_dst / b_dst: one device address
    per owner."""
    _chk(a, torch.float32, "a")
    D = a.shape[1]
    n0 = 0 if b0 is None else b0.shape[0]
    check(lib().ttam_slot_pack(a.data_ptr(), _ptr(b0), n0, _ptr(b1), req_of.data_ptr(), len(a_dst), cap, D,
                               _ptr_array(a_dst), _ptr_array(b_dst), ld_dst, _stream()), "slot_pack")


def slot_ids(src: list, cap: int, recv_idx: torch.Tensor, local_rows: torch.Tensor) -> None:
    """src: one device address per requester (its ids for this owner).  recv_idx / local_rows int64 [len(src)*cap]."""
    check(lib().ttam_slot_ids(_ptr_array(src), len(src), cap, recv_idx.data_ptr(), local_rows.data_ptr(), _stream()), "slot_ids")
