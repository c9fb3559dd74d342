"""Plans for the tcgen05 implicit-GEMM kernels (csrc/igemm.cu).

A *plan* binds fixed device buffers to one kernel launch: two TMA tensor maps
plus the tile/tap tables.  Layers build their plans once (buffers are static,
which is also what CUDA-graph capture wants) and then only call ``run()``.

Layout conventions (all bf16 unless noted):
  activations  NHWC  ``[N, H, W, C]``        C % 8 == 0
  weights      KRSC  ``[Cout, R, S, Cin]``   (a dense layer is R = S = 1)
  weight grads KRSC  fp32, accumulated with red.global.add (zero them first)

These replace the cuDNN/cuBLAS calls the reference reaches through TensorFlow
(SURVEY.md section 2.6(b)).
"""
import os

from .. import _build

BLOCK_M = 128
BLOCK_K = 64
_launches = 0
# weight-gradient tiling knobs (tools/bench_igemm.py sweeps them)
_WGRAD_WIDE = os.environ.get("TFOS_WGRAD_WIDE", "1") == "1"
_WGRAD_WIDE_MIN_PIXELS = int(os.environ.get("TFOS_WGRAD_WIDE_MIN_PIXELS", "32768"))
_WGRAD_WORK = int(os.environ.get("TFOS_WGRAD_WORK", "0"))  # 0: per-layer heuristic
_STEM_HALO = os.environ.get("TFOS_STEM_HALO", "1") == "1"
_WGRAD_HALO = os.environ.get("TFOS_WGRAD_HALO", "1") == "1"
_WGRAD_HALO_MIN_PIXELS = int(os.environ.get("TFOS_WGRAD_HALO_MIN_PIXELS", "16384"))
_STEM_SPLITS = int(os.environ.get("TFOS_STEM_SPLITS", "148"))


def launch_count():
  """Number of native kernel launches issued through this module (bench accounting)."""
  return _launches


def _C():
  return _build.load(required=True)


class Plan(object):
  """One or more prepared igemm launches; ``run()`` enqueues them on the current stream."""

  def __init__(self, handles, keep, desc=""):
    self.handles = list(handles)
    self.keep = keep  # tensors that must outlive the plan
    self.desc = desc

  def run(self):
    global _launches
    C = _C()
    for h in self.handles:
      C.igemm_run(h)
    _launches += len(self.handles)

  @property
  def num_launches(self):
    return len(self.handles)

  def set_reverse(self, flag):
    """Walk the pixel tiles last-to-first (forward-like plans): an L2 hand-over hint - start where
    the kernel that produced this plan's input stopped writing (models/resnet.py)."""
    C = _C()
    for h in self.handles:
      C.igemm_set_reverse(h, bool(flag))

  def info(self):
    return [_C().igemm_info(h) for h in self.handles]

  def __del__(self):
    try:
      C = _build.load(required=False)
      if C is not None:
        for h in self.handles:
          C.igemm_free(h)
    except Exception:
      pass
    self.handles = []


def _bn_for(n):
  if n <= 64:
    return 64
  if n <= 128:
    return 128
  return 256 if n % 256 == 0 else 128


def _tmap4(t, dims_whn, c, box, es=None, c_box=64):
  """Tensor map (C, W, H, N) over an NHWC tensor (or any tensor addressed that way)."""
  W, H, N = dims_whn
  d = {
      "base": t.data_ptr(),
      "dims": [c, W, H, N],
      "strides": [c * 2, W * c * 2, H * W * c * 2],
      "box": [c_box] + list(box),
  }
  if es is not None:
    d["elem_strides"] = [1] + list(es)
  return d


def _tmap2(t, rows, cols, box_rows, ld=None):
  """Tensor map over a row-major [rows, cols] matrix; box = 64 columns x box_rows rows."""
  ld = cols if ld is None else ld
  return {"base": t.data_ptr(), "dims": [cols, rows], "strides": [ld * 2], "box": [64, box_rows]}


def choose_box(OW, OH, N, multiple_of=1, allow_pad=False, max_rows=128, min_rows=1):
  """Pick the pixel box (bw, bh, bn) with bw*bh*bn <= max_rows that wastes the fewest MMA rows.

  ``multiple_of`` constrains the row count (the weight-gradient kernel needs a
  multiple of 16 because pixels are its K dimension); ``allow_pad`` lets the box
  overhang the tensor (TMA zero-fills the overhang).
  """
  best = None
  max_bw = min(max_rows, OW + (7 if allow_pad else 0))
  for bw in range(1, max_bw + 1):
    tw = -(-OW // bw)
    if not allow_pad and bw > OW:
      break
    for bh in range(1, min(max_rows // bw, OH + (15 if allow_pad else 0)) + 1):
      th = -(-OH // bh)
      full_img = tw == 1 and th == 1
      for bn in ([1] + ([b for b in range(2, max_rows // (bw * bh) + 1)] if full_img else [])):
        if bn > N:
          break
        rows = bw * bh * bn
        if rows % multiple_of or rows < min_rows:
          continue
        tn = -(-N // bn)
        denom = 128 if multiple_of == 1 else rows
        eff = float(OW * OH * N) / (tw * th * tn * denom)
        # prefer efficient, then wide rows (longer contiguous TMA segments)
        key = (round(eff, 4), bw, bh)
        if best is None or key > best[0]:
          best = (key, (bw, bh, bn, tw, th, tn))
  if best is None:
    raise ValueError("no box for {}x{}x{}".format(OW, OH, N))
  return best[1]


def _stats_args(g, stats):
  if stats is not None:
    g["col_sum"] = stats[0].data_ptr()
    g["col_sumsq"] = stats[1].data_ptr()


# --------------------------------------------------------------------- GEMM
def gemm(a, b, out, b_layout="nk", bias=None, relu=False, accumulate=False, stats=None):
  """out[M, N] = a[M, K] @ B (+bias)(relu).  b_layout 'nk': b is [N, K]; 'kn': b is [K, N]."""
  import torch
  M, K = a.shape
  N = b.shape[0] if b_layout == "nk" else b.shape[1]
  assert K % 8 == 0 and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
  assert a.is_contiguous() and b.is_contiguous() and out.is_contiguous()
  bn = _bn_for(N)
  ta = _tmap4(a, (M, 1, 1), K, (128, 1, 1))
  if b_layout == "nk":
    assert b.shape[1] == K
    tb = _tmap2(b, N, K, bn)
  else:
    assert b.shape[0] == K and N % 8 == 0
    tb = _tmap2(b, K, N, 64)
  g = {
      "tiles_w": -(-M // 128), "n_tiles": -(-N // bn), "box_w": 128,
      "num_taps": 1, "k_chunks": -(-K // 64),
      "lim_w": M, "lim_h": 1, "lim_n": 1, "OW": M, "OH": 1,
      "ldo": out.shape[1], "n_valid": N, "relu": int(relu),
      "out_fp32": int(out.dtype == torch.float32), "accumulate": int(accumulate),
      "bias": bias.data_ptr() if bias is not None else 0, "out": out.data_ptr(),
  }
  _stats_args(g, stats)
  h = _C().igemm_plan_fwd(ta, tb, g, bn, b_layout == "kn")
  return Plan([h], (a, b, out, bias, stats), "gemm M{} N{} K{}".format(M, N, K))


def gemm_wgrad(dy, x, dw):
  """dw[N_out, K_in] (fp32, +=) = dy[M, N_out]^T @ x[M, K_in]."""
  M, Cout = dy.shape
  Cin = x.shape[1]
  return _wgrad_plan(dy, (M, 1, 1), Cout, x, (M, 1, 1), Cin, dw, Cin, [(0, 0, 0, 0)], 1, 1,
                     box=(128, 1, 1, -(-M // 128), 1, 1))


# --------------------------------------------------------------------- conv
def conv_fprop(x, w, y, stride=1, pad=0, bias=None, relu=False, stats=None, accumulate=False):
  """y[N,OH,OW,Cout] (=|+=) conv(x[N,H,W,Cin], w[Cout,R,S,Cin]) with fused bias/ReLU/BN statistics.
  ``accumulate``: y = act(y + conv + bias) - the residual add of an inference network whose batch
  norms are folded into the filters (models/resnet.py: build_folded_inference)."""
  N, H, W, Cin = x.shape
  Cout, R, S, _ = w.shape
  _, OH, OW, _ = y.shape
  assert Cin % 8 == 0 and w.shape[3] == Cin and R * S <= 9
  bn = _bn_for(Cout)
  if R == 1 and S == 1 and stride == 1 and pad == 0:
    return gemm(x.view(-1, Cin), w.view(Cout, Cin), y.view(-1, Cout), "nk", bias, relu, accumulate,
                stats)
  bw, bh, bnn, tw, th, tn = choose_box(OW, OH, N)
  if stride == 1:
    ta = _tmap4(x, (W, H, N), Cin, (bw, bh, bnn))
  else:
    ta = _tmap4(x, (W, H, N), Cin, (bw * stride, bh * stride, bnn), es=(stride, stride, 1))
  tb = _tmap2(w, Cout, R * S * Cin, bn)
  taps = [(r, s) for r in range(R) for s in range(S)]
  g = {
      "tiles_w": tw, "tiles_h": th, "tiles_n": tn, "n_tiles": -(-Cout // bn),
      "box_w": bw, "box_h": bh, "box_n": bnn, "mul_w": stride, "mul_h": stride,
      "num_taps": len(taps), "k_chunks": -(-Cin // 64),
      "tap_dw": [s - pad for (r, s) in taps], "tap_dh": [r - pad for (r, s) in taps],
      "tap_bk": [(r * S + s) * Cin for (r, s) in taps],
      "lim_w": OW, "lim_h": OH, "lim_n": N, "OW": OW, "OH": OH,
      "ldo": Cout, "n_valid": Cout, "relu": int(relu), "accumulate": int(accumulate),
      "bias": bias.data_ptr() if bias is not None else 0, "out": y.data_ptr(),
  }
  _stats_args(g, stats)
  h = _C().igemm_plan_fwd(ta, tb, g, bn, False)
  return Plan([h], (x, w, y, bias, stats), "fprop {}x{} s{} {}->{} @{}x{}".format(
      R, S, stride, Cin, Cout, OH, OW))


def conv_dgrad(dy, w, dx, stride=1, pad=0, accumulate=False, relu=False, bias=None, stats=None,
               acc_mask=None, bn_reduce=None):
  """dx[N,H,W,Cin] (=|+=) conv_transpose(dy[N,OH,OW,Cout], w[Cout,R,S,Cin]).

  Also the forward of a transposed convolution (Keras Conv2DTranspose) when
  ``w`` is laid out [Cin_of_the_transposed_conv, R, S, Cout_of_it].
  For stride 2 the output is split in four parity classes, each a small
  stride-1 convolution over dy with the taps of matching parity.

  ``bn_reduce = (x_raw, mask_bits | None, sum_g, sum_gx)``: dx is the gradient entering the batch
  norm whose input was ``x_raw`` (shape of dx); the epilogue accumulates the two per-channel sums
  of that batch norm's backward reduction while it stores dx (csrc/igemm.h: red_x).  Requires
  that this call writes EVERY pixel of dx (raises ValueError otherwise: a strided 1x1 does not).
  """
  N, OH, OW, Cout = dy.shape
  _, R, S, Cin = w.shape
  _, H, W, _ = dx.shape
  assert Cout % 8 == 0 and Cin % 8 == 0
  bn = _bn_for(Cin)
  tb = _tmap2(w, Cout, R * S * Cin, 64)
  handles = []
  classes = [(0, 0)] if stride == 1 else [(ph, pw) for ph in range(stride) for pw in range(stride)]
  for (ph, pw) in classes:
    taps = [(r, s) for r in range(R) for s in range(S)
            if (ph + pad - r) % stride == 0 and (pw + pad - s) % stride == 0]
    if not taps:
      assert accumulate, "strided dgrad class without taps needs accumulate=True (or a pre-zeroed dx)"
      if bn_reduce is not None:
        raise ValueError("fused BN reduction: this data gradient does not write every pixel")
      continue
    ch, cw = -(-(H - ph) // stride), -(-(W - pw) // stride)  # pixels of this class
    if ch <= 0 or cw <= 0:
      continue
    bw, bh, bnn, tw, th, tn = choose_box(cw, ch, N)
    ta = _tmap4(dy, (OW, OH, N), Cout, (bw, bh, bnn))
    g = {
        "tiles_w": tw, "tiles_h": th, "tiles_n": tn, "n_tiles": -(-Cin // bn),
        "box_w": bw, "box_h": bh, "box_n": bnn,
        "num_taps": len(taps), "k_chunks": -(-Cout // 64),
        "tap_dw": [(pw + pad - s) // stride for (r, s) in taps],
        "tap_dh": [(ph + pad - r) // stride for (r, s) in taps],
        "tap_bn": [(r * S + s) * Cin for (r, s) in taps],
        "lim_w": cw, "lim_h": ch, "lim_n": N, "OW": W, "OH": H,
        "osw": stride, "oow": pw, "osh": stride, "ooh": ph,
        "ldo": Cin, "n_valid": Cin, "accumulate": int(accumulate), "relu": int(relu),
        "acc_mask": acc_mask.data_ptr() if acc_mask is not None else 0,
        "bias": bias.data_ptr() if bias is not None else 0, "out": dx.data_ptr(),
    }
    _stats_args(g, stats)  # transposed-conv *forward* feeding a batch norm
    if bn_reduce is not None:
      rx, rmask, sum_g, sum_gx = bn_reduce
      assert stats is None and tuple(rx.shape) == tuple(dx.shape) and Cin % bn == 0
      g["red_x"] = rx.data_ptr()
      g["red_mask"] = rmask.data_ptr() if rmask is not None else 0
      g["col_sum"], g["col_sumsq"] = sum_g.data_ptr(), sum_gx.data_ptr()
    handles.append(_C().igemm_plan_fwd(ta, tb, g, bn, True))
  return Plan(handles, (dy, w, dx, bias, stats, acc_mask, bn_reduce),
              "dgrad {}x{} s{} {}->{} @{}x{}{}".format(R, S, stride, Cout, Cin, H, W,
                                                      " +bnred" if bn_reduce is not None else ""))


def _wgrad_plan(dy, dy_whn, Cout, x, x_whn, Cin, dw, ldw, taps, mul, es, box=None, c_box_b=None,
                tap_dc=None, tap_out=None, n_valid=None):
  OW, OH, N = dy_whn
  n_valid = Cin if n_valid is None else n_valid
  # 256 x 256 tile per CTA (igemm_wgrad_wide_kernel): half the L2 -> SM bytes per MAC; needs
  # 64-pixel stages, so it only pays when both channel counts fill the tile
  # (measured, tools/bench_igemm.py: 14x14 layers -15...25 %; 7x7 layers lose to box padding)
  wide = (_WGRAD_WIDE and Cout % 256 == 0 and n_valid % 256 == 0 and es == 1 and c_box_b is None
          and OW * OH * N >= _WGRAD_WIDE_MIN_PIXELS)
  if wide:
    if box is None:
      box = choose_box(OW, OH, N, multiple_of=16, allow_pad=True, max_rows=64, min_rows=64)
    elif box[0] * box[1] * box[2] > 64:   # caller's 128-row GEMM box: halve it
      assert box[1] == 1 and box[2] == 1
      box = (64, 1, 1, -(-OW // 64), 1, 1)
  if box is None:
    box = choose_box(OW, OH, N, multiple_of=16, allow_pad=True)
  bw, bh, bnn, tw, th, tn = box
  assert (bw * bh * bnn) % 16 == 0 and bw * bh * bnn <= 128
  bn = 64 if n_valid <= 64 else 128
  if wide:
    bn = 256
  ta = _tmap4(dy, dy_whn, Cout, (bw, bh, bnn))
  if es == 1:
    tb = _tmap4(x, x_whn, Cin, (bw, bh, bnn))
  else:
    tb = _tmap4(x, x_whn, Cin, (bw * es, bh * es, bnn), es=(es, es, 1))
  m_tiles, n_tiles = -(-Cout // 128), -(-n_valid // bn)
  out_tiles = len(taps) * (-(-m_tiles // 2) if wide else m_tiles) * n_tiles
  total_boxes = tw * th * tn
  # split-K work items (measured, tools/bench_igemm.py --kind wgrad): the HBM-bound 1x1 layers
  # with many pixels want one wave (fewer fp32 atomics), everything else two to four
  work = _WGRAD_WORK
  if work == 0 and wide:
    # 256 KB of fp32 atomics per work item and no epilogue overlap: one wave for the memory
    # bound 1x1 layers, two for the 3x3 ones
    work = 148 if len(taps) == 1 else 296
  if work == 0:
    if len(taps) == 1:
      work = 148 if total_boxes >= 1024 else 592
    else:
      work = 592 if Cout <= 64 else 296
  k_splits = max(1, min(-(-work // out_tiles), max(1, total_boxes // 2)))
  g = {
      "tiles_w": tw, "tiles_h": th, "tiles_n": tn, "box_w": bw, "box_h": bh, "box_n": bnn,
      "mul_w": mul, "mul_h": mul, "num_taps": len(taps),
      "tap_dw": [t[0] for t in taps], "tap_dh": [t[1] for t in taps],
      "tap_dc": tap_dc if tap_dc is not None else [0] * len(taps),
      "tap_out": tap_out if tap_out is not None else [t[2] for t in taps],
      "m_tiles": m_tiles, "n_tiles": n_tiles, "k_splits": k_splits,
      "m_valid": Cout, "n_valid": n_valid, "ldw": ldw, "dw": dw.data_ptr(), "wide": int(wide),
  }
  h = _C().igemm_plan_wgrad(ta, tb, g, bn)
  return Plan([h], (dy, x, dw), "wgrad{} {}->{} taps{} splits{}".format(
      " wide" if wide else "", Cin, Cout, len(taps), k_splits))


def _wgrad3x3_halo_plan(dy, x, dw):
  """3x3 / stride 1 / pad 1 with Cin = 64, Cout <= 64: eight taps in one pass over the pixels
  (csrc/igemm_wgrad.cu igemm_wgrad_halo_kernel), the centre tap by the generic kernel."""
  N, OH, OW, Cout = dy.shape
  _, H, W, Cin = x.shape
  ta = _tmap4(dy, (OW, OH, N), Cout, (8, 16, 1))
  tb = _tmap4(x, (W, H, N), Cin, (8, 16 + 2, 1))
  taps = [(dh, dw_) for dh in (-1, 0, 1) for dw_ in (-1, 0, 1) if (dh, dw_) != (0, 0)]
  tw, th = -(-OW // 8), -(-OH // 16)
  g = {
      "tiles_w": tw, "tiles_h": th, "tiles_n": N, "box_w": 8, "box_h": 16, "box_n": 1,
      "num_taps": 1, "stem": 2,
      "halo_boxes": 3, "box_dw": [-1, 0, 1], "halo_rows": 16 + 2, "halo_hmul": 1, "halo_h0": -1,
      "halo_jmul": 2, "halo_rowbytes": 1024, "halo_nacc": 8,
      "acc_box": [d + 1 for (_, d) in taps], "acc_row": [dh + 1 for (dh, _) in taps],
      "tap_out": [((dh + 1) * 3 + (d + 1)) * Cin for (dh, d) in taps],
      "m_tiles": 1, "n_tiles": 1, "k_splits": max(1, min(_STEM_SPLITS, tw * th * N)),
      "m_valid": Cout, "n_valid": 64, "ldw": 9 * Cin, "dw": dw.data_ptr(),
  }
  halo = _C().igemm_plan_wgrad(ta, tb, g, 64)
  centre = _wgrad_plan(dy, (OW, OH, N), Cout, x, (W, H, N), Cin, dw, 9 * Cin, [(0, 0, 4 * Cin)], 1, 1)
  handles, centre.handles = [halo] + centre.handles, []   # this plan owns (and frees) them all
  return Plan(handles, (dy, x, dw), "wgrad 3x3 halo {}->{}".format(Cin, Cout))


def conv_wgrad(dy, x, dw, stride=1, pad=0):
  """dw[Cout,R,S,Cin] (fp32, +=) = sum_pixels dy[.,Cout]^T x_shifted[.,Cin]."""
  N, OH, OW, Cout = dy.shape
  _, H, W, Cin = x.shape
  _, R, S, _ = dw.shape
  if R == 1 and S == 1 and stride == 1 and pad == 0:
    return gemm_wgrad(dy.view(-1, Cout), x.view(-1, Cin), dw.view(Cout, Cin))
  if (_WGRAD_HALO and R == 3 and S == 3 and stride == 1 and pad == 1 and Cin == 64 and Cout <= 64
      and Cout % 8 == 0 and N * OH * OW >= _WGRAD_HALO_MIN_PIXELS):
    return _wgrad3x3_halo_plan(dy, x, dw)
  taps = [(s - pad, r - pad, (r * S + s) * Cin) for r in range(R) for s in range(S)]
  return _wgrad_plan(dy, (OW, OH, N), Cout, x, (W, H, N), Cin, dw, R * S * Cin, taps, stride,
                     stride)


# --------------------------------------------------------------- ResNet stem
# 7x7 / stride-2 convolution on a 3-channel image.  The input is stored as
# [N, H, Wp, 8] (3 real + 5 zero channels, zero columns left and right) so that
# one filter ROW - 7 taps x 8 channels = 56 values, padded to 64 - is a single
# contiguous 128-byte TMA row; an overlapping tensor map (W stride = 2 pixels)
# turns the whole 7x7 window into 7 K-blocks of 64.
STEM_K, STEM_PAD, STEM_CP = 7, 3, 8


def stem_geometry(H, W):
  OH, OW = (H + 2 * STEM_PAD - STEM_K) // 2 + 1, (W + 2 * STEM_PAD - STEM_K) // 2 + 1
  Wp = 2 * (OW - 1) + 8 + 0  # last window start + 8 pixels
  Wp = max(Wp, W + STEM_PAD + 1)
  Wp = (Wp + 1) // 2 * 2
  return OH, OW, Wp


def _stem_tmap(xp, N, H, Wp, OW, box, es_h):
  bw, bh, bnn = box
  return {
      "base": xp.data_ptr(),
      "dims": [64, OW, H, N],
      "strides": [2 * STEM_CP * 2, Wp * STEM_CP * 2, H * Wp * STEM_CP * 2],
      "box": [64, bw, bh * es_h, bnn],
      "elem_strides": [1, 1, es_h, 1],
  }


def _stem_halo_tmap(xp, N, H, Wp, OW, bw, rows):
  """Window-row view of xp with input rows at stride 1: one box = bw window columns x rows."""
  return {
      "base": xp.data_ptr(),
      "dims": [64, OW, H, N],
      "strides": [2 * STEM_CP * 2, Wp * STEM_CP * 2, H * Wp * STEM_CP * 2],
      "box": [64, bw, rows, 1],
  }


def stem_fprop(xp, w, y, bias=None, relu=False, stats=None):
  """xp [N,H,Wp,8] (image at column offset 3), w [Cout, 7, 64] (row-major taps, 8 ch each), y [N,OH,OW,Cout]."""
  N, H, Wp, _ = xp.shape
  Cout = w.shape[0]
  _, OH, OW, _ = y.shape
  bn = _bn_for(Cout)
  if Cout <= 64 and _STEM_HALO:
    # halo mode (csrc/igemm.cu "stem mode"): 8 x 16 output pixels per tile, one 8 x 37 box of
    # input rows serves all seven filter rows, the packed filter stays resident in smem
    ta = _stem_halo_tmap(xp, N, H, Wp, OW, 8, 2 * 16 + 5)
    tb = _tmap2(w, Cout, 7 * 64, bn)
    g = {
        "tiles_w": -(-OW // 8), "tiles_h": -(-OH // 16), "tiles_n": N, "n_tiles": 1,
        "box_w": 8, "box_h": 16, "box_n": 1, "mul_w": 1, "mul_h": 2,
        "num_taps": 7, "k_chunks": 1, "stem": 1,
        "tap_dw": [0] * 7, "tap_dh": [-STEM_PAD] * 7, "tap_bk": [r * 64 for r in range(7)],
        "lim_w": OW, "lim_h": OH, "lim_n": N, "OW": OW, "OH": OH,
        "ldo": Cout, "n_valid": Cout, "relu": int(relu),
        "bias": bias.data_ptr() if bias is not None else 0, "out": y.data_ptr(),
    }
    _stats_args(g, stats)
    h = _C().igemm_plan_fwd(ta, tb, g, bn, False)
    return Plan([h], (xp, w, y, bias, stats), "stem fprop (halo)")
  bw, bh, bnn, tw, th, tn = choose_box(OW, OH, N)
  ta = _stem_tmap(xp, N, H, Wp, OW, (bw, bh, bnn), 2)
  tb = _tmap2(w, Cout, 7 * 64, bn)
  g = {
      "tiles_w": tw, "tiles_h": th, "tiles_n": tn, "n_tiles": -(-Cout // bn),
      "box_w": bw, "box_h": bh, "box_n": bnn, "mul_w": 1, "mul_h": 2,
      "num_taps": 7, "k_chunks": 1,
      "tap_dw": [0] * 7, "tap_dh": [r - STEM_PAD for r in range(7)],
      "tap_bk": [r * 64 for r in range(7)],
      "lim_w": OW, "lim_h": OH, "lim_n": N, "OW": OW, "OH": OH,
      "ldo": Cout, "n_valid": Cout, "relu": int(relu),
      "bias": bias.data_ptr() if bias is not None else 0, "out": y.data_ptr(),
  }
  _stats_args(g, stats)
  h = _C().igemm_plan_fwd(ta, tb, g, bn, False)
  return Plan([h], (xp, w, y, bias, stats), "stem fprop")


def stem_wgrad(dy, xp, dw):
  """dw [Cout, 7, 64] fp32 += dy[N,OH,OW,Cout]^T windows(xp)."""
  N, OH, OW, Cout = dy.shape
  _, H, Wp, _ = xp.shape
  if Cout <= 64 and _STEM_HALO:
    # all seven filter rows from one 16 x 21 halo box per 16 x 8 output pixels
    # (csrc/igemm_wgrad.cu igemm_wgrad_stem_kernel)
    ta = _tmap4(dy, (OW, OH, N), Cout, (16, 8, 1))
    tb = _stem_halo_tmap(xp, N, H, Wp, OW, 16, 2 * 8 + 5)
    tw, th = -(-OW // 16), -(-OH // 8)
    g = {
        "tiles_w": tw, "tiles_h": th, "tiles_n": N, "box_w": 16, "box_h": 8, "box_n": 1,
        "num_taps": 1, "stem": 1,
        "halo_boxes": 1, "box_dw": [0], "halo_rows": 2 * 8 + 5, "halo_hmul": 2,
        "halo_h0": -STEM_PAD, "halo_jmul": 2, "halo_rowbytes": 2048, "halo_nacc": 7,
        "acc_box": [0] * 7, "acc_row": list(range(7)), "tap_out": [r * 64 for r in range(7)],
        "m_tiles": 1, "n_tiles": 1, "k_splits": max(1, min(_STEM_SPLITS, tw * th * N)),
        "m_valid": Cout, "n_valid": 64, "ldw": 7 * 64, "dw": dw.data_ptr(),
    }
    h = _C().igemm_plan_wgrad(ta, tb, g, 64)
    return Plan([h], (dy, xp, dw), "stem wgrad (halo)")
  bw, bh, bnn, tw, th, tn = choose_box(OW, OH, N, multiple_of=16, allow_pad=True)
  ta = _tmap4(dy, (OW, OH, N), Cout, (bw, bh, bnn))
  tb = _stem_tmap(xp, N, H, Wp, OW, (bw, bh, bnn), 2)
  out_tiles = 7
  k_splits = max(1, min(-(-296 // out_tiles), max(1, tw * th * tn // 2)))
  g = {
      "tiles_w": tw, "tiles_h": th, "tiles_n": tn, "box_w": bw, "box_h": bh, "box_n": bnn,
      "mul_w": 1, "mul_h": 2, "num_taps": 7,
      "tap_dw": [0] * 7, "tap_dh": [r - STEM_PAD for r in range(7)], "tap_dc": [0] * 7,
      "tap_out": [r * 64 for r in range(7)],
      "m_tiles": -(-Cout // 128), "n_tiles": 1, "k_splits": k_splits,
      "m_valid": Cout, "n_valid": 64, "ldw": 7 * 64, "dw": dw.data_ptr(),
  }
  h = _C().igemm_plan_wgrad(ta, tb, g, 64)
  return Plan([h], (dy, xp, dw), "stem wgrad")


def pack_stem_weight(w_krsc):
  """[Cout, 7, 7, 3] float -> [Cout, 7, 64] (7 taps x 8 channels + 8 zeros per filter row)."""
  import torch
  Cout = w_krsc.shape[0]
  out = torch.zeros(Cout, 7, 64, dtype=w_krsc.dtype, device=w_krsc.device)
  out[:, :, :56].view(Cout, 7, 7, 8)[..., :3] = w_krsc
  return out


def unpack_stem_weight(w_packed):
  Cout = w_packed.shape[0]
  return w_packed[:, :, :56].reshape(Cout, 7, 7, 8)[..., :3].contiguous()
