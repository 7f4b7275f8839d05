"""Stand-alone timing of the training step's individual kernels at ConvNeXt-B shapes (not the bench contract).

    python tools/prof_train_kernels.py [stage] [batch] [iters]        # CUDA-event timings, one JSON line per kernel
    ncu --set full ... python tools/prof_train_kernels.py 2 128 1     # the same launches for an ncu capture

Stage s of ConvNeXt-B: (H, C) = (56,128) (28,256) (14,512) (7,1024).
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from visiondk_b200 import _lib

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
only = sys.argv[4].split(",") if len(sys.argv) > 4 else None
H = (56, 28, 14, 7)[stage]
Cn = (128, 256, 512, 1024)[stage]
M = B * H * H
lib = _lib.load()
dev = "cuda"
sp = _lib.stream_ptr()


def bf(*shape, scale=1.0):
    return (scale * torch.randn(*shape, device=dev)).to(torch.bfloat16)


x = bf(B, H, H, Cn)
y = torch.empty_like(x)
dy = bf(B, H, H, Cn)
rstd = torch.rand(M, device=dev) + 0.5
w49 = 0.2 * torch.randn(49, Cn, device=dev)
vec = lambda: torch.randn(Cn, device=dev)
bias, ln_w, ln_b = vec(), vec() + 2.0, vec()
dw49 = torch.zeros(49, Cn, device=dev)
dbias, dgamma, dbeta = torch.zeros(Cn, device=dev), torch.zeros(Cn, device=dev), torch.zeros(Cn, device=dev)
hpre = torch.empty(M, 4 * Cn, dtype=torch.bfloat16, device=dev)
hpost = torch.empty(M, 4 * Cn, dtype=torch.bfloat16, device=dev)
w1 = bf(4 * Cn, Cn, scale=0.05)
w2 = bf(Cn, 4 * Cn, scale=0.05)
b1 = torch.randn(4 * Cn, device=dev)
gam = torch.randn(Cn, device=dev)
out = torch.empty_like(x)
slabs = torch.empty(64 * 1024 * 1024 // 4, device=dev)


def gemm(A, Bm, D, M_, N_, K_, lda, ldb, ldd, epi=_lib.EPI_NONE, bias=0, gamma=0, residual=0, ldr=0, out_dtype=_lib.DTYPE_BF16,
         split=1, stride=0, aux=0, ta=0, tb=0):
    g = _lib.GemmDesc(A=A.data_ptr(), B=Bm.data_ptr(), D=D.data_ptr(), M=M_, N=N_, K=K_, lda=lda, ldb=ldb, ldd=ldd,
                      in_dtype=_lib.DTYPE_BF16, out_dtype=out_dtype, epilogue=epi, bias=bias, gamma=gamma, beta=0, residual=residual,
                      ldr=ldr, ln_eps=1e-6, split_k=split, split_stride=stride, aux_out=aux, trans_a=ta, trans_b=tb)
    _lib.check(lib.vdk_gemm(C.byref(g), sp), "gemm")


xf = x.reshape(M, Cn)
dxf = dy.reshape(M, Cn)
tiles = ((Cn + 127) // 128) * ((4 * Cn + 255) // 256)
split = lib.vdk_gemm_effective_splits(M, max(2, 296 // tiles))

kernels = {
    "dwconv7_ln fwd": (lambda: _lib.check(lib.vdk_dwconv7(0, x.data_ptr(), B, H, H, Cn, w49.data_ptr(), bias.data_ptr(), ln_w.data_ptr(),
                                                          ln_b.data_ptr(), 1e-6, y.data_ptr(), rstd.data_ptr(), 0, sp), "dw0"),
                       2.0 * M * Cn * 49, 4.0 * M * Cn),
    "dwconv7 bwd-data": (lambda: _lib.check(lib.vdk_dwconv7(1, dy.data_ptr(), B, H, H, Cn, w49.data_ptr(), 0, 0, 0, 0.0, y.data_ptr(), 0,
                                                            x.data_ptr(), sp), "dw1"), 2.0 * M * Cn * 49, 6.0 * M * Cn),
    "dwconv7 wgrad": (lambda: _lib.check(lib.vdk_dwconv7_wgrad(x.data_ptr(), dy.data_ptr(), B, H, H, Cn, dw49.data_ptr(), dbias.data_ptr(), sp),
                                         "dww"), 2.0 * M * Cn * 49, 4.0 * M * Cn),
    "ln_bwd": (lambda: _lib.check(lib.vdk_layernorm_bwd(dy.data_ptr(), x.data_ptr(), rstd.data_ptr(), B, H, H, Cn, ln_w.data_ptr(),
                                                        ln_b.data_ptr(), 1, y.data_ptr(), 0, dgamma.data_ptr(), dbeta.data_ptr(), sp), "lnb"),
               0.0, 6.0 * M * Cn),
    "fc1 fwd (GELU + saved pre-activation)": (lambda: gemm(xf, w1, hpost, M, 4 * Cn, Cn, Cn, Cn, 4 * Cn, _lib.EPI_GELU, b1.data_ptr(),
                                                           aux=hpre.data_ptr()), 8.0 * M * Cn * Cn, 2.0 * M * Cn * 9),
    "fc1 fwd (GELU only)": (lambda: gemm(xf, w1, hpost, M, 4 * Cn, Cn, Cn, Cn, 4 * Cn, _lib.EPI_GELU, b1.data_ptr()),
                            8.0 * M * Cn * Cn, 2.0 * M * Cn * 5),
    "fc1 fwd (bias only)": (lambda: gemm(xf, w1, hpost, M, 4 * Cn, Cn, Cn, Cn, 4 * Cn, _lib.EPI_NONE, b1.data_ptr()),
                            8.0 * M * Cn * Cn, 2.0 * M * Cn * 5),
    "fc1 fwd (no epilogue math)": (lambda: gemm(xf, w1, hpost, M, 4 * Cn, Cn, Cn, Cn, 4 * Cn, _lib.EPI_NONE),
                                   8.0 * M * Cn * Cn, 2.0 * M * Cn * 5),
    "fc2 fwd (layer scale + residual)": (lambda: gemm(hpost, w2, out, M, Cn, 4 * Cn, 4 * Cn, 4 * Cn, Cn, _lib.EPI_SCALE_RESIDUAL,
                                                      bias.data_ptr(), gam.data_ptr(), xf.data_ptr(), Cn), 8.0 * M * Cn * Cn, 2.0 * M * Cn * 6),
    "fc2 dgrad (x gelu')": (lambda: gemm(dxf, w2, hpost, M, 4 * Cn, Cn, Cn, 4 * Cn, 4 * Cn, _lib.EPI_MUL_GELU_GRAD, residual=hpre.data_ptr(),
                                         ldr=4 * Cn, tb=1), 8.0 * M * Cn * Cn, 2.0 * M * Cn * 9),
    "fc1 dgrad": (lambda: gemm(hpost, w1, y.reshape(M, Cn), M, Cn, 4 * Cn, 4 * Cn, Cn, Cn, tb=1), 8.0 * M * Cn * Cn, 2.0 * M * Cn * 5),
    "fc2 wgrad (slabs)": (lambda: gemm(dxf, hpost, slabs, Cn, 4 * Cn, M, Cn, 4 * Cn, 4 * Cn, out_dtype=_lib.DTYPE_FP32, split=split,
                                       stride=Cn * 4 * Cn, ta=1, tb=1), 8.0 * M * Cn * Cn, 2.0 * M * Cn * 5),
}

for name, (fn, flops, bytes_) in kernels.items():
    if only and not any(o in name for o in only):
        continue
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(json.dumps({"kernel": name, "stage": stage, "batch": B, "us": round(us, 1), "tflops": round(flops / us / 1e6, 1),
                      "algo_GBps": round(bytes_ / us / 1e3, 1)}))
