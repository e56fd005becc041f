#!/usr/bin/env python
"""Where a training-GEMM CTA spends its cycles: clock64 stamps of CTA (0,0,0) of gemm_tc_kernel (train_tc.cu) taken at
kernel start, after the prologue (barrier init + TMEM alloc), after griddepcontrol.wait, at the first / last K-block the MMA
warp sees, when the epilogue gets the accumulator, after the epilogue and after TMEM dealloc -- printed relative to the
first stamp.  (K-block period = (stamp[4] - stamp[3]) / (K/64 - 1); round 1: 480 cycles at a 128x64 tile against a
128-cycle MMA floor: per-SM operand ingest ~60 B/clk and the dependent tcgen05.mma chain ~120 cycles each.)

    python benchmarks/gemm_tc_stamps.py
"""
import ctypes as C, torch, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddqst_b200 as dq
lib = dq._lib.load()
f = lib.ddqst_selftest_gemm_tc_dbg
f.restype = C.c_int
f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
for (m, n, k) in [(1024, 512, 512), (128, 64, 64), (128, 64, 1024)]:
    a = torch.randn(m, k, device='cuda').to(torch.bfloat16)
    b = torch.randn(n, k, device='cuda').to(torch.bfloat16)
    c = torch.empty(m, n, device='cuda')
    dbg = torch.zeros(8, dtype=torch.int64, device='cuda')
    for it in range(3):
        dq._lib.check(f(a.data_ptr(), b.data_ptr(), 0, 0, m, n, k, 1, c.data_ptr(), dbg.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        d = dbg.cpu().tolist()
        print((m, n, k), it, [x - d[0] for x in d])
