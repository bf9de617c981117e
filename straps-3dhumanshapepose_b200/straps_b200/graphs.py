"""CUDA-graph capture of the hot path (inference).

A forward of the drop-in modules is a fixed sequence of ~27 library kernels plus a handful of tensor allocations; at
B=64 the launch gaps are 2.3 % of the step (`tools/graph_probe.py`: 1.74 ms eager, 1.70 ms replayed; bit-identical), and
proportionally more at small batch.  `GraphedCallable` captures any function of CUDA tensors built from those modules
once (fixed shapes) and replays it: inputs are copied into the captured input buffers, outputs are the captured output
tensors (valid until the next call).  Nothing on the path synchronises or allocates outside torch's caching allocator,
and the library's internal fork/join of the downsample convolutions is a capturable pattern.
"""
import torch

from ._lib import StrapsError


def _tensors(x):
    if torch.is_tensor(x):
        return [x]
    if isinstance(x, (tuple, list)):
        return [t for item in x for t in _tensors(item)]
    return []


class GraphedCallable(object):
    def __init__(self, fn, *example_inputs, warmup=3):
        for t in example_inputs:
            if not torch.is_tensor(t) or not t.is_cuda:
                raise StrapsError('GraphedCallable captures functions of CUDA tensors (got %r)' % (type(t),))
        self._fn = fn
        self._static_in = [t.detach().clone() for t in example_inputs]
        dev = self._static_in[0].device if self._static_in else torch.device('cuda', torch.cuda.current_device())
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):          # lazy weight packing, tensor maps, allocator warm-up happen here
                fn(*self._static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph), torch.no_grad():
            self._static_out = fn(*self._static_in)
        if not _tensors(self._static_out):
            raise StrapsError('GraphedCallable: the captured function returned no tensors')

    @property
    def static_inputs(self):
        """The captured input buffers: a caller that writes its inputs straight into them (e.g. as the destination of its host->device
        copy) and passes them back skips the copy `__call__` otherwise makes."""
        return list(self._static_in)

    def __call__(self, *inputs):
        if len(inputs) != len(self._static_in):
            raise StrapsError('GraphedCallable was captured with %d inputs, called with %d' % (len(self._static_in), len(inputs)))
        for dst, src in zip(self._static_in, inputs):
            if src.shape != dst.shape or src.dtype != dst.dtype:
                raise StrapsError('GraphedCallable: input %s/%s does not match the captured %s/%s'
                                  % (tuple(src.shape), src.dtype, tuple(dst.shape), dst.dtype))
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._static_out


class GraphedTrainStep(object):
    """One whole training step -- forward, loss, backward, gradient exchange, Adam -- captured as CUDA graph(s) and replayed.

    `step_fn(*inputs)` must run the step exactly as the reference's loop body does (train/train_synthetic_otf_rendering.py:186-233):
    `optimiser.zero_grad(); loss, ... = criterion(...); loss.backward(); optimiser.step()` with a
    straps_b200.parallel.DataParallelAdam `optimiser`, and return the tensors the caller wants to read DETACHED (e.g.
    `loss.detach()`: a returned tensor that still holds its autograd graph keeps the graph's nodes and their streams alive across
    steps).  Joint selections inside the step use `straps_b200.ops.select_joints` (indexing with a Python list copies the list to the
    device on every call, which a capture refuses).  The step
    launches ~500 kernels for ~9 ms of GPU work; eager PyTorch spends another ~1.5 ms in launch gaps and autograd bookkeeping,
    which replay removes.  Everything the step touches is graph-safe: the library allocates its workspaces on first use (the
    warm-up steps here), Adam's step count lives on the device, BatchNorm's counters are bumped by a captured foreach kernel.

    world_size 1: one graph.  world_size > 1: the all-reduce stays OUTSIDE the capture (NCCL work is enqueued on the replaying stream
    between the two graphs): graph A = zero_grad + forward + backward + gather, collective, graph B = Adam.  The split is done by
    the optimiser: `step_fn` is captured with `optimiser.defer_exchange()` active.
    """

    def __init__(self, step_fn, optimiser, *example_inputs, warmup=3):
        from .parallel import DataParallelAdam
        if not isinstance(optimiser, DataParallelAdam):
            raise StrapsError('GraphedTrainStep needs a straps_b200.parallel.DataParallelAdam (device-side step count, flat buckets)')
        for t in example_inputs:
            if not torch.is_tensor(t) or not t.is_cuda:
                raise StrapsError('GraphedTrainStep captures functions of CUDA tensors (got %r)' % (type(t),))
        self._opt = optimiser
        self._static_in = [t.detach().clone() for t in example_inputs]
        dev = optimiser.bucket.params.device
        import gc
        gc.collect()                                 # autograd graphs of earlier eager steps (and their AccumulateGrad nodes) go first
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(2, warmup)):          # workspaces, tensor maps, allocator warm-up; these are REAL training steps
                step_fn(*self._static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._split = optimiser.world > 1
        optimiser._deferred = self._split
        # capture on the warm-up stream: autograd nodes that outlive a step (AccumulateGrad of the criterion's log-variances) were
        # created there, and a node on ANOTHER stream makes the engine synchronise with it, which invalidates the capture
        try:
            try:
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph, stream=side):
                    self._static_out = step_fn(*self._static_in)
            except Exception:
                # an early all-reduce inside the backward pass (DataParallelAdam.enable_overlap) is an NCCL collective inside the
                # capture; if this build of torch / NCCL refuses that, capture again with the whole collective between the graphs
                if optimiser._early_range is None:
                    raise
                optimiser.disable_overlap()
                torch.cuda.synchronize(dev)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph, stream=side):
                    self._static_out = step_fn(*self._static_in)
        finally:
            optimiser._deferred = False
        self._graph_b = None
        if self._split:
            self._graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_b, pool=self._graph.pool(), stream=side):
                optimiser.apply_update()
            # the capture ran the forward/backward once without the update: finish that step so the replicas stay in step
            optimiser.exchange()
            self._graph_b.replay()

    def __call__(self, *inputs):
        for dst, src in zip(self._static_in, inputs):
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        if self._split:
            self._opt.exchange()
            self._graph_b.replay()
        self._opt.bucket.bump_versions()             # host-side bookkeeping the replay skipped: the packed weight copies are stale
        return self._static_out
