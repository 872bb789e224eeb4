"""PCIe ceiling for the host-buffer path: one-way and simultaneous two-way copy rates of pinned buffers, and
the rhs_host / step_2n_host times over the chunk-size option."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import trixi_b200 as T  # noqa: E402


def copies():
    n = 5368709120 // 8
    a = torch.empty(n, dtype=torch.float64).pin_memory()
    b = torch.empty(n, dtype=torch.float64).pin_memory()
    da = torch.empty(n, dtype=torch.float64, device="cuda")
    db = torch.zeros(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}

    def t(fn):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def h2d():
        with torch.cuda.stream(s1):
            da.copy_(a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            b.copy_(db, non_blocking=True)

    def both():
        h2d()
        d2h()

    out["h2d_gbs"] = n * 8 / t(h2d) / 1e9
    out["d2h_gbs"] = n * 8 / t(d2h) / 1e9
    tb = t(both)
    out["both_s"] = tb
    out["both_gbs_each"] = n * 8 / tb / 1e9
    return out


def main():
    res = {"copies": copies()}
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    semi = bench.make_semi(level, device=0)
    gpu = semi.backend()
    n = semi.u_length()
    u = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    du = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    u0 = T.compute_coefficients(0.0, semi).ravel(order="F")
    alg = T.CarpenterKennedy2N54()
    gpu.set_option(gpu.OPT_FUSED_CFL, 1)
    for chunk in [0, -1, 512, 1024, 4096, 8192, 32768]:
        gpu.set_option(gpu.OPT_HOST_PIPELINE_CHUNK, chunk)
        u[:] = u0
        gpu.rhs_host(du, u, 0.0)
        t0 = time.perf_counter()
        for _ in range(2):
            gpu.rhs_host(du, u, 0.0)
        t_rhs = (time.perf_counter() - t0) / 2
        gpu.upload(0, u)
        dt = 1.3 * gpu.max_dt()
        gpu.step_2n_host(u, 0.0, dt, alg.a, alg.b, alg.c)
        t0 = time.perf_counter()
        for _ in range(2):
            gpu.step_2n_host(u, 0.0, dt, alg.a, alg.b, alg.c)
        t_step = (time.perf_counter() - t0) / 2
        res[f"chunk_{chunk}"] = {"rhs_host_ms": t_rhs * 1e3, "step_2n_host_ms": t_step * 1e3}
        print(chunk, res[f"chunk_{chunk}"], file=sys.stderr, flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
