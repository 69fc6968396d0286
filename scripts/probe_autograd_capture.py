"""Experiment: does an autograd.grad w.r.t. leaf parameters inside a CUDA-graph capture fail when the
parameters' AccumulateGrad nodes were created on the legacy stream (held alive by an outer graph)?"""
import torch
dev = "cuda"
lin = torch.nn.Linear(8, 8).to(dev)
x = torch.randn(4, 8, device=dev)
s = torch.cuda.Stream()


def inner():
    xx = x.detach().requires_grad_(True)
    with torch.enable_grad():
        f = lin(xx)
        return torch.autograd.grad(f, (xx,) + tuple(lin.parameters()), torch.ones_like(f))


def attempt(tag, outer_on_side):
    if outer_on_side:
        with torch.cuda.stream(s):
            outer = sum(p.sum() for p in lin.parameters())      # keeps the accumulators alive
    else:
        outer = sum(p.sum() for p in lin.parameters())
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        inner()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, stream=s):
            inner()
        print(tag, "capture OK")
    except Exception as e:
        print(tag, "capture FAILED:", str(e).splitlines()[0])
    torch.cuda.synchronize()
    del outer


attempt("outer graph built on legacy stream:", False)
attempt("outer graph built on capture stream:", True)
attempt("outer graph built on legacy stream (again):", False)
