"""e2e (host buffers in/out) clips/s for different gs_segment_host chunk counts."""
import os, subprocess, sys
for n in (1, 2, 3, 4):
    env = dict(os.environ, GOLFER_HOST_CHUNKS=str(n))
    code = ("import sys,time,torch;sys.path.insert(0,'.');import golfer_b200;"
            "seg=golfer_b200.Segmenter(golfer_b200.V0,precision='bf16',max_B=256,max_T=300);"
            "x=torch.randn(256,300,17,3).pin_memory();"
            "[seg.segment(x) for _ in range(3)];torch.cuda.synchronize();t0=time.perf_counter();"
            "[seg.segment(x) for _ in range(10)];torch.cuda.synchronize();dt=time.perf_counter()-t0;"
            "print('chunks',%d,'e2e clips/s',round(2560/dt))" % n)
    subprocess.run([sys.executable, "-c", code], env=env)
