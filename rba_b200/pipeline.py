"""Overlapped host->device->host scoring stream (pinned-memory prefetch and true batching).

The reference's evaluation loop is batch-1 and synchronous: per image a blocking H2D, the forward, and a blocking
`.cpu().numpy()` (support.py:353-399, evaluate_ood.py:143-150).  `ScoreStream` keeps the same per-batch semantics
(uint8 CHW images in host memory -> float32 score maps in host memory) but runs the three legs on three CUDA streams:

    copy-in stream   H2D of batch i+1 (pinned staging)       |
    compute stream   rba_forward of batch i (CUDA graph)     |  all concurrent
    copy-out stream  D2H of the scores of batch i-1          |

so at steady state a step costs max(forward, H2D, D2H) instead of their sum.  Results come back in order, one step
late; `close()`/exhausting the iterator drains the pipeline.  torch is only the carrier of buffers, streams and events.
"""
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from ._lib import RbaError


class ScoreStream:
    """eng: rba_b200.Engine with weights loaded.  B, H, W: fixed batch shape (the forward is captured once)."""

    def __init__(self, eng, B, H, W, use_graph=True, post_forward=None, d2h=True):
        if not torch.cuda.is_available():
            raise RbaError("ScoreStream needs a CUDA device: the product path has no CPU fallback")
        self.eng, self.B, self.H, self.W = eng, B, H, W
        self.post_forward = post_forward      # callable(static_out) launched on the compute stream after each forward
        self.d2h = d2h                        # False: scores stay on the device (step() returns this step's device tensor)
        dev = eng.device
        self.dev = dev
        proto = torch.empty((B, 3, H, W), dtype=torch.uint8, device=dev)
        self.use_graph = use_graph
        if use_graph:
            self.static_in, self.static_out, self.graph = eng.graphed(proto, rba=True)
            self._gen = eng.arena_generation()
        else:
            eng.reserve(B, H, W)
            self.static_in = proto
            self.static_out = eng.alloc_outputs(B, H, W, rba=True)
            self.graph = None
        self.stage_in = [torch.empty_like(proto) for _ in range(2)]
        self.stage_out = [torch.empty((B, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
        self.host_out = [torch.empty((B, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ev = lambda: torch.cuda.Event()  # noqa: E731
        self.ev_in_ready = [ev(), ev()]      # H2D into stage_in[k] finished
        self.ev_in_free = [ev(), ev()]       # compute has consumed stage_in[k]
        self.ev_c_done = [ev(), ev()]        # scores of this step are in stage_out[k]
        self.ev_out_done = [ev(), ev()]      # D2H out of stage_out[k] finished
        self.n_submitted = 0
        self.n_collected = 0
        self.h2d_bytes_per_step = proto.numel()
        self.d2h_bytes_per_step = B * H * W * 4

    # ---- one step = submit batch i, (maybe) collect batch i-1 ----
    def submit(self, host_images):
        """host_images: (B,3,H,W) uint8 tensor in PINNED host memory.  Asynchronous."""
        if host_images.shape != self.stage_in[0].shape or host_images.dtype != torch.uint8:
            raise RbaError(f"ScoreStream.submit: expected uint8 {tuple(self.stage_in[0].shape)}, got {host_images.dtype} {tuple(host_images.shape)}")
        if not host_images.is_pinned():
            raise RbaError("ScoreStream.submit: host batch must be in pinned memory (torch.Tensor.pin_memory())")
        i = self.n_submitted
        k = i & 1
        comp = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.s_in):
            if i >= 2:
                self.s_in.wait_event(self.ev_in_free[k])
            self.stage_in[k].copy_(host_images, non_blocking=True)
            self.ev_in_ready[k].record(self.s_in)
        comp.wait_event(self.ev_in_ready[k])
        if self.graph is not None and self.eng.arena_generation() != self._gen:
            # the engine's arena was outgrown by a larger forward elsewhere: re-capture in the new arena (the old graph
            # stays valid -- its arena is retired, not freed -- but would pin that memory)
            self.static_in, self.static_out, self.graph = self.eng.graphed(self.static_in, rba=True)
            self._gen = self.eng.arena_generation()
        self.static_in.copy_(self.stage_in[k], non_blocking=True)
        self.ev_in_free[k].record(comp)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.eng.forward_into(self.static_in, self.static_out)
        if self.post_forward is not None:
            self.post_forward(self.static_out)
        self.n_submitted += 1
        if not self.d2h:
            self.n_collected = self.n_submitted
            return
        if i >= 2:
            comp.wait_event(self.ev_out_done[k])          # the D2H that last read stage_out[k]
        self.stage_out[k].copy_(self.static_out["rba"], non_blocking=True)
        self.ev_c_done[k].record(comp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_c_done[k])
            self.host_out[k].copy_(self.stage_out[k], non_blocking=True)
            self.ev_out_done[k].record(self.s_out)

    def collect(self):
        """Blocks until the oldest un-collected batch's scores are in host memory; returns the (B,H,W) pinned tensor
        (valid until two more batches have been submitted)."""
        if self.n_collected >= self.n_submitted:
            raise RbaError("ScoreStream.collect: nothing in flight")
        k = self.n_collected & 1
        self.ev_out_done[k].synchronize()
        self.n_collected += 1
        return self.host_out[k]

    def step(self, host_images):
        """Submit batch i; return the scores of batch i-1 (None on the first call).  With d2h=False: returns the DEVICE
        score tensor of batch i itself (stream-ordered; valid until the next submit)."""
        if not self.d2h:
            self.submit(host_images)
            return self.static_out["rba"]
        prev = None
        if self.n_submitted - self.n_collected >= 2:
            raise RbaError("ScoreStream.step: collect() the previous result first")
        self.submit(host_images)
        if self.n_submitted - self.n_collected == 2:
            prev = self.collect()
        return prev

    def drain(self):
        out = []
        while self.n_collected < self.n_submitted:
            out.append(self.collect())
        return out

    def run(self, host_batches):
        """Generator: yields one (B,H,W) host score tensor per input batch, in order."""
        for hb in host_batches:
            r = self.step(hb)
            if r is not None:
                yield r
        for r in self.drain():
            yield r


class PinnedBatcher:
    """Host side of the input pipeline.  The reference decodes one image at a time on the main thread
    (`DataLoader(batch_size=1)` + albumentations `ToTensorV2`, support.py:73-81,353-372) and its scorers read only `x[0]`
    (evaluate_ood.py:146).  Here `dataset[i]` (any indexable returning `(image, label)` with image CHW/HWC uint8 and label
    HW, tensors or ndarrays -- e.g. the reference's dataset classes) is called from `workers` threads, and consecutive
    samples are packed into fixed-shape batches of `batch` images in PINNED buffers (a ring of `depth` buffer sets) ready
    for `ScoreStream.submit`.  A short last batch is padded by repeating its first image with label 255 (ignored by
    the metrics, support.py:275-279).  Iterating yields `(images (B,3,H,W) uint8, labels (B,H,W) uint8, n_valid)`."""

    def __init__(self, dataset, batch, indices=None, workers=8, depth=3, pin=None):
        self.ds, self.B = dataset, int(batch)
        self.idx = list(range(len(dataset))) if indices is None else list(indices)
        self.workers, self.depth = max(1, int(workers)), max(2, int(depth))
        self.pin = torch.cuda.is_available() if pin is None else bool(pin)
        self._bufs = None

    def __len__(self):
        return -(-len(self.idx) // self.B)

    @staticmethod
    def _chw_u8(img):
        t = torch.as_tensor(np.ascontiguousarray(img)) if not torch.is_tensor(img) else img
        if t.dim() != 3:
            raise RbaError(f"PinnedBatcher: image must have 3 dims, got {tuple(t.shape)}")
        if t.shape[0] != 3 and t.shape[-1] == 3:
            t = t.permute(2, 0, 1)                        # HWC -> CHW (what ToTensorV2 does)
        if t.dtype != torch.uint8:
            raise RbaError(f"PinnedBatcher: uint8 images expected (raw pixel values), got {t.dtype}")
        return t

    def _alloc(self, H, W):
        mk = (lambda *s: torch.empty(s, dtype=torch.uint8).pin_memory()) if self.pin else (lambda *s: torch.empty(s, dtype=torch.uint8))
        self._bufs = [(mk(self.B, 3, H, W), mk(self.B, H, W)) for _ in range(self.depth)]
        self._free = queue.Queue()
        for b in self._bufs:
            self._free.put(b)

    def release(self, images):
        """Hand a yielded buffer set back to the ring (the default iteration does it when the next batch is requested)."""
        for b in self._bufs:
            if b[0] is images:
                self._free.put(b)
                return

    def __iter__(self):
        out = queue.Queue(maxsize=self.depth)
        stop = threading.Event()

        def load(i):
            img, lab = self.ds[i]
            lab = torch.as_tensor(np.ascontiguousarray(lab)) if not torch.is_tensor(lab) else lab
            # only 0 (in-distribution) and 1 (OoD) count (support.py:275-279); anything else -- including int64 values
            # that would wrap to 0/1 in a uint8 cast -- becomes the ignore value 255
            lab = torch.where((lab == 0) | (lab == 1), lab, torch.full_like(lab, 255)).to(torch.uint8)
            return self._chw_u8(img), lab

        def put(item):                            # polls `stop`: a consumer that left cannot strand the producer
            while not stop.is_set():
                try:
                    out.put(item, timeout=0.2)
                    return
                except queue.Full:
                    pass

        def producer():
            try:
                with ThreadPoolExecutor(self.workers) as ex:
                    for s in range(0, len(self.idx), self.B):
                        if stop.is_set():
                            return
                        chunk = self.idx[s:s + self.B]
                        samples = list(ex.map(load, chunk))
                        H, W = samples[0][0].shape[-2:]
                        if self._bufs is None:
                            self._alloc(H, W)
                        for im, lab in samples:
                            if tuple(im.shape[-2:]) != (H, W) or tuple(lab.shape[-2:]) != (H, W) or \
                                    tuple(self._bufs[0][0].shape[-2:]) != (H, W):
                                raise RbaError("PinnedBatcher: all images of a run must share one size "
                                               f"(got {tuple(im.shape[-2:])} vs {(H, W)}); resize in the dataset transform")
                        while True:               # polls `stop` so an abandoned iteration cannot strand this thread
                            try:
                                ib, lb = self._free.get(timeout=0.2)
                                break
                            except queue.Empty:
                                if stop.is_set():
                                    return
                        for j, (im, lab) in enumerate(samples):
                            ib[j].copy_(im)
                            lb[j].copy_(lab.reshape(H, W))
                        for j in range(len(samples), self.B):   # pad: repeated image, ignored labels
                            ib[j].copy_(ib[0])
                            lb[j].fill_(255)
                        put((ib, lb, len(samples)))
                put(None)
            except BaseException as e:   # surface loader errors in the consumer
                put(e)

        t = threading.Thread(target=producer, daemon=True)
        t.start()
        prev = None
        try:
            while True:
                item = out.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                if prev is not None:
                    self.release(prev)
                prev = item[0]
                yield item
        finally:
            stop.set()
            if prev is not None and self._bufs is not None:
                self.release(prev)


# ---------------------------------------------------------------------------------------------------------
# Device-side JPEG decode: bitstreams -> planar RGB uint8 in HBM
# ---------------------------------------------------------------------------------------------------------
def jpeg_available():
    """True when the nvJPEG run-time library could be bound (rba_jpeg_available)."""
    from . import _lib
    return bool(_lib.lib().rba_jpeg_available())


def jpeg_info(data):
    """(height, width, channels) of one JPEG bitstream (bytes / bytearray / uint8 array)."""
    import ctypes
    from . import _lib
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    h, w, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.lib().rba_jpeg_info(buf.ctypes.data, buf.size, ctypes.byref(h), ctypes.byref(w), ctypes.byref(c)))
    return h.value, w.value, c.value


def decode_jpeg_batch(streams, device=None, out=None):
    """streams: sequence of JPEG bitstreams (bytes or uint8 arrays, or paths of .jpg files) of ONE image size.
    Returns a (n,3,H,W) uint8 CUDA tensor of planar RGB -- the layout `Engine.forward` / `ScoreStream.submit` take --
    decoded on the device by nvJPEG (`rba_jpeg_decode`), stream-ordered on the current CUDA stream.  Replaces
    `cv2.imread` / `PIL.Image.open` + `ToTensorV2` of the reference's dataset classes (support.py:73-81) for JPEG inputs."""
    import ctypes
    from . import _lib
    if not torch.cuda.is_available():
        raise RbaError("decode_jpeg_batch needs a CUDA device: the product path has no CPU fallback")
    bufs = []
    for s in streams:
        if isinstance(s, str):
            with open(s, "rb") as f:
                s = f.read()
        bufs.append(np.frombuffer(s, dtype=np.uint8) if not isinstance(s, np.ndarray) else np.ascontiguousarray(s, dtype=np.uint8))
    n = len(bufs)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if n == 0:
        return torch.empty((0, 3, 0, 0), dtype=torch.uint8, device=dev)
    H, W, _ = jpeg_info(bufs[0])
    if out is None:
        out = torch.empty((n, 3, H, W), dtype=torch.uint8, device=dev)
    elif tuple(out.shape) != (n, 3, H, W) or out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous():
        raise RbaError(f"decode_jpeg_batch: out must be a contiguous uint8 CUDA tensor of shape {(n, 3, H, W)}")
    ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
    sizes = (ctypes.c_int64 * n)(*[b.size for b in bufs])
    with torch.cuda.device(out.device):
        st = torch.cuda.current_stream(out.device).cuda_stream
        _lib.check(_lib.lib().rba_jpeg_decode(ptrs, sizes, n, out.data_ptr(), H, W, ctypes.c_void_p(st)))
    return out
