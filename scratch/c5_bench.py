"""C5 (arm_ball n=27 m=7 N=400 B=512 setInterval-5): per-phase times of the all-active iterations for
a few candidates-per-round settings (scratch tool)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from drake_ddp_b200 import problems
prob = problems.arm_ball(400)
x0 = prob.batch_x0(512, seed=0)
for A in [int(a) for a in (sys.argv[1:] or ["8", "18", "27"])]:
    r = bench.solve_batch_throughput(torch, prob, 512, x0, A=A)
    print(A, json.dumps(r["all_active"]), r["ms_per_batch_iteration"], r["status"])
