"""CUDA-graph replay of a whole hot-path step (forward + autograd backward [+ optimizer]) built from the layer API.

The fused kernels make a step a handful of launches, so through the reference's nn.Module API the step is bound by what
the HOST does per batch -- ~80 per-feature Parameters cross autograd (one AccumulateGrad node each) although two kernels do
all the work.  `GraphedStep` captures one eager execution of the user's step function into a CUDA graph and replays it:
the launches (ours, the side-stream zero-fill of the gradient buffer, torch's own glue) are re-issued by the driver with
no Python in between.  The usual CUDA-graph contract applies: the step reads its inputs from STATIC tensors (copy each
batch into them, e.g. PackedBatch.copy_into), shapes are fixed, and the tensors it returns / the .grad tensors it leaves
are static as well (overwritten by every replay)."""
import torch


class GraphedStep(object):
    def __init__(self, fn, warmup=3, stream=None, pool=None):
        """fn(): one step on static inputs; returns a tensor / tuple of tensors (or None).  Parameters' .grad must be
        None (or be reset by fn itself) when fn runs, so that the backward allocates them from the graph's pool."""
        self.fn = fn
        dev = torch.cuda.current_device()
        side = stream or torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):            # warm-up off the default stream: lazy inits, plan caches, allocator
            for _ in range(max(1, warmup)):
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool, stream=side):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out

    def pool(self):
        return self.graph.pool()
