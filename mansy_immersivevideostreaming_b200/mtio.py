"""MTIO viewport-prediction transformer, inference on the GPU through the C ABI.

Host-side mirror of the reference's ``ViewportTransformerMTIO`` (viewport_prediction/models/mtio.py:48-133) for
the calls ``predict.py`` makes on it: construct (predict.py:68-75), ``load_state_dict`` (predict.py:17),
``eval`` / ``to`` (predict.py:22,92) and ``sample(history, current)`` (predict.py:27).  A checkpoint written by
``run_models.py`` (the model's own state dict) loads unchanged; keys that a modern torch omits
(``*.bias`` of the transformer, see oracle/mtio_oracle.py) are taken as zeros.  Training (``forward``,
``loss_function``) is out of scope.

``predict_chunk_masks`` is the rest of ``predict.predict`` (predict.py:33-48): the first ``dataset_frequency``
predicted points of every sample -> tile masks + IoU, on the device (``mansy_viewport_tiles``), i.e. the
``vp_pred`` column of the simulator's tables (BASELINE config 5).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Mapping, Optional, Tuple

import numpy as np
import torch

from . import _capi
from ._capi import MansyError, MtioWeights, check
from .simulator import ViewportTiler, _require_cuda

D_MODEL = 512
TOKEN = 6          # in_channel 2 x 3 MTIO heads


def _arr(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def positional_encoding(n: int, d: int = D_MODEL) -> np.ndarray:
    """The registered buffer of ``PositionalEncoding`` (mtio.py:18-26) when the state dict does not carry it."""
    pe = np.zeros((n, d), dtype=np.float32)
    position = np.arange(0, n, dtype=np.float32)[:, None]
    div_term = np.exp(np.arange(0, d, 2, dtype=np.float32) * np.float32(-(math.log(10000.0) / d))).astype(np.float32)
    pe[:, 0::2] = np.sin(position * div_term)
    pe[:, 1::2] = np.cos(position * div_term)
    return pe


# ---- deterministic weights and synthetic inputs (numpy only: identical in the build container and on the GPU box) ----
def mtio_state_dict_shapes(n_enc: int = 2, n_dec: int = 2, d: int = D_MODEL, ff: int = D_MODEL, bias: bool = True):
    s = [("embedding.linear.weight", (d, TOKEN)), ("embedding.linear.bias", (d,))]

    def attn(p):
        out = [(p + "in_proj_weight", (3 * d, d))]
        if bias:
            out.append((p + "in_proj_bias", (3 * d,)))
        out.append((p + "out_proj.weight", (d, d)))
        if bias:
            out.append((p + "out_proj.bias", (d,)))
        return out

    def lin(p, o, i):
        return [(p + ".weight", (o, i))] + ([(p + ".bias", (o,))] if bias else [])

    def norm(p):
        return [(p + ".weight", (d,))] + ([(p + ".bias", (d,))] if bias else [])

    for l in range(n_enc):
        p = f"transformer.encoder.layers.{l}."
        s += attn(p + "self_attn.") + lin(p + "linear1", ff, d) + lin(p + "linear2", d, ff) + norm(p + "norm1") + norm(p + "norm2")
    s += norm("transformer.encoder.norm")
    for l in range(n_dec):
        p = f"transformer.decoder.layers.{l}."
        s += (attn(p + "self_attn.") + attn(p + "multihead_attn.") + lin(p + "linear1", ff, d) + lin(p + "linear2", d, ff)
              + norm(p + "norm1") + norm(p + "norm2") + norm(p + "norm3"))
    s += norm("transformer.decoder.norm")
    s += [("transformer.distill_layer.downConv.weight", (d, d, 3)), ("transformer.distill_layer.downConv.bias", (d,)),
          ("transformer.distill_layer.norm.weight", (d,)), ("transformer.distill_layer.norm.bias", (d,)),
          ("transformer.distill_layer.norm.running_mean", (d,)), ("transformer.distill_layer.norm.running_var", (d,)),
          ("predictor.0.weight", (TOKEN, d)), ("predictor.0.bias", (TOKEN,))]
    return s


def seeded_mtio_state_dict(seed: int, bias: bool = True, n_enc: int = 2, n_dec: int = 2) -> dict:
    """Matrices uniform(+-1/sqrt(fan_in)); norm gains 1 + 0.1u, norm / linear biases 0.1u; BatchNorm running
    statistics away from (0, 1) so that the eval-mode affine is exercised."""
    rng = np.random.default_rng(seed)
    sd: dict = {}
    for name, shape in mtio_state_dict_shapes(n_enc, n_dec, bias=bias):
        u = rng.uniform(-1.0, 1.0, size=shape)
        if len(shape) > 1:
            v = u / math.sqrt(float(np.prod(shape[1:])))
        elif name.endswith("running_var"):
            v = 1.0 + 0.5 * np.abs(u)
        elif "norm" in name and name.endswith("weight"):
            v = 1.0 + 0.1 * u
        else:
            v = 0.1 * u
        sd[name] = v.astype(np.float32)
    return sd


def synthetic_history(n: int, seed: int, his_window: int = 5):
    """5 Hz random walk of viewport centres on the unit torus: history [n,M,2], current [n,1,2]."""
    rng = np.random.default_rng(seed)
    start = rng.uniform(0.05, 0.95, size=(n, 1, 2))
    steps = rng.normal(0.0, 0.03, size=(n, his_window + 1, 2))
    walk = np.mod(start + np.cumsum(steps, axis=1), 1.0).astype(np.float32)
    return walk[:, :his_window].copy(), walk[:, his_window:].copy()



class ViewportTransformerMTIO:
    """Same constructor arguments as the reference class (mtio.py:48-50); only the configuration the reference's
    own scripts build is supported: ``in_channel=2``, ``num_head=3``, ``d_model = dim_feedforward = 512``."""

    def __init__(self, in_channel: int = 2, fut_window: int = 15, d_model: int = D_MODEL, dim_feedforward: int = D_MODEL,
                 num_head: int = 3, num_encoder_layers: int = 2, num_decoder_layers: int = 2, batch_first: bool = True,
                 dropout: float = 0.2, device="cuda", repeat_prob: float = 0.5, seed: int = 1, his_window: int = 5,
                 max_batch: int = 16384):
        if in_channel != 2 or num_head != 3 or d_model != D_MODEL or dim_feedforward != D_MODEL or not batch_first:
            raise MansyError("supported MTIO configuration: in_channel=2, num_head=3, d_model=dim_feedforward=512, batch_first")
        dev = torch.device(device)
        index = 0 if dev.index is None else dev.index
        if dev.type != "cuda":
            raise MansyError("the B200 MTIO path has no CPU fallback (device must be cuda)")
        _require_cuda(index)
        self.lib = _capi.load_library()
        self.device = torch.device("cuda", index)
        self.fut_window, self.his_window = int(fut_window), int(his_window)
        self.n_enc, self.n_dec = int(num_encoder_layers), int(num_decoder_layers)
        self.max_batch = int(max_batch)
        self.fp32 = False                   # True: exact-fp32 CUDA-core GEMMs instead of tcgen05 TF32
        self._h: Optional[C.c_void_p] = None

    # -- nn.Module surface predict.py touches --
    def to(self, device):
        if torch.device(device).type != "cuda":
            raise MansyError("the B200 MTIO path has no CPU fallback")
        return self

    def eval(self):
        return self

    def load_state_dict(self, sd: Mapping[str, object], strict: bool = True):
        keep = []
        w = MtioWeights()
        w.n_enc, w.n_dec, w.his_window, w.fut_window = self.n_enc, self.n_dec, self.his_window, self.fut_window

        def ptr(key: str, shape, required: bool = True):
            if key not in sd:
                if required:
                    raise KeyError(key)
                return None
            a = _arr(sd[key])
            if a.size != int(np.prod(shape)):
                raise ValueError(f"{key}: expected shape {tuple(shape)}, got {a.shape}")
            keep.append(a)
            return a.ctypes.data

        d = D_MODEL
        w.emb_w, w.emb_b = ptr("embedding.linear.weight", (d, TOKEN)), ptr("embedding.linear.bias", (d,))
        rows = max(self.his_window, self.fut_window) + 1
        if "positional_embedding.pe" in sd:
            pe = _arr(sd["positional_embedding.pe"]).reshape(-1, d)[:rows].copy()
        else:
            pe = positional_encoding(rows)
        keep.append(pe)
        w.pe, w.pe_rows = pe.ctypes.data, pe.shape[0]

        def attn(dst, p):
            dst.in_proj_w = ptr(p + "in_proj_weight", (3 * d, d))
            dst.in_proj_b = ptr(p + "in_proj_bias", (3 * d,), False)
            dst.out_w = ptr(p + "out_proj.weight", (d, d))
            dst.out_b = ptr(p + "out_proj.bias", (d,), False)

        for side, n in (("encoder", self.n_enc), ("decoder", self.n_dec)):
            for l in range(n):
                p = f"transformer.{side}.layers.{l}."
                lay = (w.enc if side == "encoder" else w.dec)[l]
                attn(lay.self_attn, p + "self_attn.")
                if side == "decoder":
                    attn(lay.cross_attn, p + "multihead_attn.")
                    lay.norm3_w, lay.norm3_b = ptr(p + "norm3.weight", (d,)), ptr(p + "norm3.bias", (d,), False)
                lay.lin1_w, lay.lin1_b = ptr(p + "linear1.weight", (d, d)), ptr(p + "linear1.bias", (d,), False)
                lay.lin2_w, lay.lin2_b = ptr(p + "linear2.weight", (d, d)), ptr(p + "linear2.bias", (d,), False)
                lay.norm1_w, lay.norm1_b = ptr(p + "norm1.weight", (d,)), ptr(p + "norm1.bias", (d,), False)
                lay.norm2_w, lay.norm2_b = ptr(p + "norm2.weight", (d,)), ptr(p + "norm2.bias", (d,), False)
            if strict and f"transformer.{side}.layers.{n}.linear1.weight" in sd:
                raise ValueError(f"state dict has more than {n} {side} layers")
        w.enc_norm_w, w.enc_norm_b = ptr("transformer.encoder.norm.weight", (d,)), ptr("transformer.encoder.norm.bias", (d,), False)
        w.dec_norm_w, w.dec_norm_b = ptr("transformer.decoder.norm.weight", (d,)), ptr("transformer.decoder.norm.bias", (d,), False)
        p = "transformer.distill_layer."
        w.conv_w, w.conv_b = ptr(p + "downConv.weight", (d, d, 3)), ptr(p + "downConv.bias", (d,))
        w.bn_w, w.bn_b = ptr(p + "norm.weight", (d,)), ptr(p + "norm.bias", (d,))
        w.bn_mean, w.bn_var = ptr(p + "norm.running_mean", (d,)), ptr(p + "norm.running_var", (d,))
        w.pred_w, w.pred_b = ptr("predictor.0.weight", (TOKEN, d)), ptr("predictor.0.bias", (TOKEN,))
        self.close()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.mansy_mtio_create(C.byref(w), self.device.index, self.max_batch, C.byref(h)))
        self._h = h
        return self

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.mansy_mtio_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _handle(self):
        if self._h is None:
            raise MansyError("load_state_dict() first: the model has no weights")
        return self._h

    def sample(self, history, current, return_tokens: bool = False, timed: bool = False, steps: int = 0):
        """history ``[B, his_window, 2]``, current ``[B, 1, 2]`` -> ensembled, wrapped viewports ``[B, fut_window, 2]``
        (mtio.py:106-133).  CUDA tensors run in place on the current stream; numpy / CPU tensors go through the
        host-buffer entry point (copies + synchronise) and come back as the input's kind.  ``steps`` (0 = all) stops
        the autoregression early; later rows of the result are then zero."""
        h = self._handle()
        flags = (_capi.MTIO_FP32 if self.fp32 else 0) | (_capi.MTIO_TIME_KERNELS if timed else 0)
        on_device = isinstance(history, torch.Tensor) and history.is_cuda
        if on_device:
            hist = history.to(device=self.device, dtype=torch.float32).contiguous()
            cur = current.to(device=self.device, dtype=torch.float32).contiguous()
            n = hist.shape[0]
            if hist.shape[1:] != (self.his_window, 2) or cur.shape != (n, 1, 2):
                raise ValueError("history must be [B, his_window, 2] and current [B, 1, 2]")
            alloc = torch.zeros if steps else torch.empty
            pred = alloc((n, self.fut_window, 2), dtype=torch.float32, device=self.device)
            tokens = alloc((n, self.fut_window + 1, TOKEN), dtype=torch.float32, device=self.device) if return_tokens else None
            check(self.lib.mansy_mtio_sample(h, hist.data_ptr(), cur.data_ptr(), n, int(steps), flags, pred.data_ptr(),
                                             None if tokens is None else tokens.data_ptr(), self._stream()))
            return (pred, tokens) if return_tokens else pred
        as_numpy = not isinstance(history, torch.Tensor)
        hist, cur = _arr(history), _arr(current)
        n = hist.shape[0]
        if hist.shape[1:] != (self.his_window, 2) or cur.shape != (n, 1, 2):
            raise ValueError("history must be [B, his_window, 2] and current [B, 1, 2]")
        pred = np.zeros((n, self.fut_window, 2), dtype=np.float32)
        check(self.lib.mansy_mtio_sample_host(h, hist.ctypes.data, cur.ctypes.data, n, int(steps), flags & ~_capi.MTIO_TIME_KERNELS,
                                              pred.ctypes.data, self._stream()))
        return pred if as_numpy else torch.from_numpy(pred)

    def sample_host(self, history: torch.Tensor, current: torch.Tensor, out: torch.Tensor, steps: int = 0) -> torch.Tensor:
        """Pinned host tensors in / out (the end-to-end path of bench.py)."""
        n = history.shape[0]
        flags = _capi.MTIO_FP32 if self.fp32 else 0
        check(self.lib.mansy_mtio_sample_host(self._handle(), history.data_ptr(), current.data_ptr(), n, int(steps), flags, out.data_ptr(),
                                              self._stream()))
        return out

    def kernel_ms(self) -> Tuple[np.ndarray, np.ndarray]:
        """After a ``sample(..., timed=True)`` and a synchronise: (ms, launches) of [GEMM, attention, other] kernels."""
        ms, cnt = (C.c_double * 3)(), (C.c_int32 * 3)()
        check(self.lib.mansy_mtio_kernel_ms(self._handle(), C.byref(ms), C.byref(cnt)))
        return np.array(ms[:]), np.array(cnt[:])

    def predict_chunk_masks(self, history, current, gt_future, tiler: Optional[ViewportTiler] = None, frequency: int = 5,
                            short: bool = True):
        """predict.py:27-48 on the device: sample, then OR the tile masks of the first ``frequency`` ground-truth and
        predicted points and take their IoU.  ``gt_future [B, >= frequency, 2]`` -> (gt masks int64 [B], predicted
        masks int64 [B], IoU float64 [B], predictions [B, fut_window, 2]).  ``short`` (default) stops the autoregression
        after the ``frequency`` steps the masks use (rows beyond are zero); a prediction never depends on later steps."""
        tiler = tiler or ViewportTiler(device=self.device.index)
        pred = self.sample(history.to(self.device), current.to(self.device), steps=min(frequency, self.fut_window) if short else 0)
        gt = gt_future.to(device=self.device, dtype=torch.float32)[:, :frequency].contiguous()
        gt_m, pred_m, acc = tiler.chunk_masks_device(gt, pred[:, :frequency].contiguous())
        return gt_m, pred_m, acc, pred


class LinearRegression:
    """The ``--model regression`` predictor of predict.py (viewport_prediction/models/linear_regression.py): same
    constructor and ``sample(history, current)``; a least-squares line per sample and coordinate on the GPU instead of
    two sklearn fits per sample in a Python loop."""

    def __init__(self, fut_window: int, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise MansyError("the B200 path has no CPU fallback (device must be cuda)")
        _require_cuda(0 if dev.index is None else dev.index)
        self.lib = _capi.load_library()
        self.device = torch.device("cuda", 0 if dev.index is None else dev.index)
        self.fut_window = int(fut_window)

    def to(self, device):
        return self

    def eval(self):
        return self

    def sample(self, history, current) -> torch.Tensor:
        hist = torch.as_tensor(history).to(device=self.device, dtype=torch.float32).contiguous()
        cur = torch.as_tensor(current).to(device=self.device, dtype=torch.float32).contiguous()
        n, his = hist.shape[0], hist.shape[1]
        if hist.shape[2:] != (2,) or cur.shape != (n, 1, 2):
            raise ValueError("history must be [B, his_window, 2] and current [B, 1, 2]")
        pred = torch.empty((n, self.fut_window, 2), dtype=torch.float32, device=self.device)
        check(self.lib.mansy_linreg_sample(hist.data_ptr(), cur.data_ptr(), n, his, self.fut_window, pred.data_ptr(),
                                           torch.cuda.current_stream(self.device).cuda_stream))
        return pred


def viewport_windows(gt_xy: np.ndarray, his_window: int = 5) -> Tuple[np.ndarray, np.ndarray]:
    """Cut the samples ``predict.py`` feeds the model out of continuous 5 Hz viewport traces.

    ``gt_xy [P, CV, F, 2]`` holds, per (video, user) pair, the ``F`` ground-truth points of each of ``CV`` consecutive
    chunks.  ``ViewportDataset.__getitem__`` (viewport_prediction/utils/load_dataset.py:48-57) with
    ``sample_step = frequency`` makes the sample of a chunk end right before it: ``current`` = the last point of the
    previous chunk, ``history`` = the ``his_window`` points before that.  The points a real trace has ahead of the
    first chunk (``trim_head``, config.yml:147) are not part of the synthetic tables; the first point is repeated
    there.  Returns ``history [P*CV, his_window, 2]``, ``current [P*CV, 1, 2]``.
    """
    P, CV, F, _ = gt_xy.shape
    pts = np.ascontiguousarray(gt_xy, dtype=np.float32).reshape(P, CV * F, 2)
    full = np.concatenate([np.repeat(pts[:, :1], his_window + 1, axis=1), pts], axis=1)
    idx = np.arange(CV) * F                               # chunk i starts at full[:, his_window + 1 + F * i]
    hist = np.stack([full[:, idx + k] for k in range(his_window)], axis=2)
    cur = full[:, idx + his_window][:, :, None, :]
    return hist.reshape(P * CV, his_window, 2).copy(), cur.reshape(P * CV, 1, 2).copy()


class MtioMaskFn:
    """``mask_fn`` of ``synth.make_synthetic_tables`` whose predicted-viewport column comes from the MTIO model
    instead of a noise model: BASELINE config 5 (MTIO inference feeding predicted viewports into the environments).
    Ground-truth masks, predicted masks and IoU are those of predict.py:33-48."""

    def __init__(self, net: ViewportTransformerMTIO, n_vp_chunks: int, tiler: Optional[ViewportTiler] = None):
        self.net, self.cv = net, int(n_vp_chunks)
        self.tiler = tiler or ViewportTiler(device=net.device.index)
        self.last_pred: Optional[np.ndarray] = None

    def __call__(self, gt_xy: np.ndarray, pred_xy: np.ndarray):
        n, F, _ = gt_xy.shape
        hist, cur = viewport_windows(np.asarray(gt_xy).reshape(n // self.cv, self.cv, F, 2), self.net.his_window)
        gt_m, pred_m, acc, pred = self.net.predict_chunk_masks(torch.from_numpy(hist), torch.from_numpy(cur),
                                                               torch.from_numpy(np.ascontiguousarray(gt_xy)), self.tiler, frequency=F)
        self.last_pred = pred[:, :F].cpu().numpy()
        return gt_m.cpu().numpy().view(np.uint64), pred_m.cpu().numpy().view(np.uint64), acc.cpu().numpy()


def write_prediction_files(results_dir: str, video: int, user: int, first_chunk: int, gt_masks, pred_masks, accuracy) -> str:
    """Write one (video, user) pair's per-chunk viewports in the reference's on-disk formats (predict.py:50-65):
    ``video{v}/user{u}.pkl`` -- the list of ``(chunk, gt uint8[64], pred uint8[64], accuracy float64)`` tuples that
    ``simulators/hmdtrace.py:5-23`` (and ``tables.pack_from_reference_layout``) load -- and the ``.csv`` twin.
    ``gt_masks`` / ``pred_masks`` are 64-bit tile masks (bit t = tile t) as the kernels produce them."""
    import os
    import pickle

    from .tables import u64_to_masks
    def bits(x):
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        if isinstance(x, np.ndarray) and x.dtype == np.int64:       # the kernels' masks come back as int64 tensors
            return x.reshape(-1).view(np.uint64)
        return np.array(x, dtype=np.uint64).reshape(-1)             # python ints: no detour through float64

    gt = u64_to_masks(bits(gt_masks))
    pred = u64_to_masks(bits(pred_masks))
    acc = np.asarray(accuracy, dtype=np.float64).reshape(-1)
    value = [(int(first_chunk + i), gt[i].reshape(-1).astype(np.uint8), pred[i].reshape(-1).astype(np.uint8), np.float64(acc[i]))
             for i in range(len(acc))]
    base = os.path.join(results_dir, f"video{video}")
    os.makedirs(base, exist_ok=True)
    path = os.path.join(base, f"user{user}.pkl")
    with open(path, "wb") as fh:
        pickle.dump(value, fh)
    with open(os.path.join(base, f"user{user}.csv"), "w", encoding="utf-8") as fh:
        fh.write("chunk,gt,pred,accuracy\n")
        for chunk, g, p, a in value:
            fh.write(f"{chunk},{','.join(map(str, list(g)))},{','.join(map(str, list(p)))},{a}\n")
    return path
