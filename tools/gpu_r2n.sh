#!/bin/bash
mkdir -p gpurun_out
BENCH_DUMP_KERNELS=gpurun_out/r2n_voc_kernels.txt timeout 600 python bench.py --steps 20 --warmup 5 --no-config5 --no-cpu-baseline --no-gpu-eager --min-seconds 0.5 > gpurun_out/r2n_bench.log 2> gpurun_out/r2n_bench.err; echo "rc=$?"
cat gpurun_out/r2n_voc_kernels.txt | head -20
# eager per-layer times (one stream, warm L2): act vs conv per stage
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
import megatts2_hierspeechpp_b200 as hsv
from megatts2_hierspeechpp_b200 import ops
dev="cuda:0"
def gt(fn, n=20):
    fn(); torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    best=1e9
    for _ in range(3):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best=min(best,e0.elapsed_time(e1)*1e3/n)
    return best
for (C,L) in ((128,1000),(64,2000),(256,2000),(128,10000),(64,40000),(32,80000),(16,160000)):
    x=torch.randn(1,C,L,device=dev); a=torch.zeros(C,device=dev)
    buf=ops.blk16_buffer(1,C,L,dev)
    ta=gt(lambda: ops.act1d_blk16(x,a,a,buf))
    line=f"C={C:4d} L={L:7d} act {ta:6.2f} us"
    for k in (3,7,11):
        w=torch.randn(C,C,k,device=dev)*0.05
        nt=ops.pick_n_tile(C, (L+127)//128, C*k)
        wp=ops.pack_conv_weight(w,nt)
        out=torch.empty_like(x)
        tc=gt(lambda: ops.conv1d_umma(buf,wp,a,L,C,C,k,1,nt,residual=x,out=out))
        line+=f"  conv k{k} {tc:6.2f}"
    print(line, flush=True)
PY
