"""torch.library registration (unidet3d_b200/torch_ops.py): schemas + fake (meta) implementations on CPU; the real
implementations against unidet3d_b200.ops on the GPU."""
import pytest
import torch


def test_ops_are_registered_with_fake_impls():
    import unidet3d_b200.torch_ops  # noqa: F401
    ns = torch.ops.unidet3d_b200
    for name in ("gather_gemm", "act_split", "segmented_mean", "layernorm", "attention", "voxel_mean"):
        assert hasattr(ns, name), name
    m = "meta"
    x = torch.empty((1000, 64), device=m)
    assert ns.gather_gemm(x, torch.empty(10, dtype=torch.uint8, device=m), 27, 64, 96, 777).shape == (777, 96)
    assert ns.act_split(torch.empty((50, 6), device=m)).shape == (50, 32)
    assert ns.segmented_mean(x, torch.empty(1000, dtype=torch.int64, device=m), 33).shape == (33, 64)
    assert ns.segmented_mean(x, torch.empty(1000, dtype=torch.int64, device=m), 33, None, 3).shape == (33, 3)
    assert ns.layernorm(x, torch.empty(64, device=m), torch.empty(64, device=m)).shape == (1000, 64)
    assert ns.attention(torch.empty((500, 768), device=m), torch.empty(3, dtype=torch.int32, device=m), 300, 8).shape == (500, 256)
    assert ns.voxel_mean(torch.empty((100, 6), device=m), torch.empty(100, dtype=torch.int32, device=m), 40).shape == (40, 6)


def test_no_cpu_path_behind_the_custom_ops():
    import unidet3d_b200.torch_ops  # noqa: F401
    from unidet3d_b200 import _lib
    with pytest.raises((_lib.Ud3dError, RuntimeError)):
        torch.ops.unidet3d_b200.layernorm(torch.zeros(4, 32), torch.ones(32), torch.zeros(32))


@pytest.mark.gpu
def test_custom_ops_match_the_ops_module_and_capture_in_a_cuda_graph():
    import unidet3d_b200.torch_ops  # noqa: F401
    from unidet3d_b200 import ops
    ns = torch.ops.unidet3d_b200
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(3000, 64, device="cuda", generator=g)
    w = torch.randn(96, 1, 64, device="cuda", generator=g) * 0.1
    pw = ops.PackedWeight(w)
    b = torch.randn(96, device="cuda", generator=g)
    ref = ops.gemm(x, pw, bias=b, act="relu")
    out = ns.gather_gemm(x, pw.data, 1, 64, 96, 3000, None, None, b, None, "relu")
    assert torch.equal(out, ref)
    gam, bet = torch.rand(64, device="cuda", generator=g) + 0.5, torch.randn(64, device="cuda", generator=g)
    assert torch.equal(ns.layernorm(x, gam, bet), ops.layernorm(x, gam, bet))
    xs = ns.act_split(x, gam, bet, True)
    assert torch.equal(xs, ops.act_split(x, gam, bet, relu=True))
    # capturable: no host synchronisation inside (allocations come from the graph's private pool)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            ns.gather_gemm(x, pw.data, 1, 64, 96, 3000, None, None, b, None, "relu")
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y = ns.layernorm(ns.gather_gemm(x, pw.data, 1, 64, 96, 3000, None, None, b, None, "relu"),
                         torch.ones(96, device="cuda"), torch.zeros(96, device="cuda"))
    x.add_(1.0)
    graph.replay()
    torch.cuda.synchronize()
    want = ops.layernorm(ops.gemm(x, pw, bias=b, act="relu"), torch.ones(96, device="cuda"), torch.zeros(96, device="cuda"))
    assert torch.equal(y, want)
