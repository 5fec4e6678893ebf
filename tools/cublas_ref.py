"""cuBLAS(Lt) on the same shapes: what the reference's GPU stage (cuda_server.c:468-491, four
cublasLtMatmul calls, CUBLAS_COMPUTE_32F) does when simply recompiled for B200, plus the TF32
library rate on a large square GEMM (calibrates the tensor roofline for 4-byte operands).
Run on the GPU box; prints one JSON line."""
import json
import sys

import torch


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dims = {"small": [352, 1024, 512, 256, 1], "medium": [880, 1024, 512, 256, 1], "large": [3968, 2048, 512, 256, 1]}
    model = sys.argv[1] if len(sys.argv) > 1 else "small"
    d = dims[model]
    out = {"model": model}
    dev = torch.device("cuda")
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        key = "tf32" if tf32 else "fp32"
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        ms = timed(lambda: torch.matmul(a, b), reps=10)
        out[f"square8192_{key}_tflops"] = 2 * n ** 3 / ms / 1e9
        del a, b
        for B in (2048, 16384):
            x = torch.randn(B, d[0], device=dev)
            W = [torch.randn(d[k], d[k + 1], device=dev) / d[k] ** 0.5 for k in range(4)]

            def chain():
                h = x
                for k in range(4):
                    h = torch.matmul(h, W[k])       # the reference's LINEAR chain, no bias/activation
                return h
            per = [timed(lambda k=k, h=torch.randn(B, d[k], device=dev): torch.matmul(h, W[k])) for k in range(4)]
            ms = timed(chain)
            out[f"{key}_B{B}"] = {"chain_ms": ms, "layer_ms": per, "inferences_per_s": B / ms * 1e3,
                                  "tflops": 2 * B * sum(d[k] * d[k + 1] for k in range(4)) / ms / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
