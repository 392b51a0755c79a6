"""Profiling target: S3 warmed into the contact regime, then a few ticks between cudaProfilerStart/Stop.

  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:<pattern> -c <n> \
      -o gpurun_out/<name> python scripts/profile_tick.py [--warm 70] [--ticks 1] [--bodies 20834]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--warm", type=int, default=70)
    ap.add_argument("--ticks", type=int, default=1)
    ap.add_argument("--bodies", type=int, default=20834)
    args = ap.parse_args()
    import torch
    import pies_b200 as pb
    from pies_b200 import scenes
    s = pb.Solver(device=0, **scenes.S3_OPTIONS)
    scenes.build_s3(s, args.bodies)
    for _ in range(args.warm):
        s.tick()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.ticks):
        s.tick()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    st = s.stats()
    print("tick %d: %.2f ms, pcg %d, pt %d floor %d, launches %d" % (
        args.warm + args.ticks, st.msTick, st.pcgIterationsLastTick, st.triCollisions, st.staticCollisions,
        st.kernelLaunchesLastTick))


if __name__ == "__main__":
    main()
