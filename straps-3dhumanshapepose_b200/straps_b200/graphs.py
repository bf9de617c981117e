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
