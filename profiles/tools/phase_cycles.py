"""Where do the critical chains of the cfg2 extend spend their cycles?  Needs the profiling build:
    make -C manipulation_planning_suite_b200/csrc libmps_b200_prof.so
    MPS_B200_LIB=manipulation_planning_suite_b200/csrc/libmps_b200_prof.so python profiles/tools/phase_cycles.py
(clock64() reads perturb the kernel a little: use for proportions, not absolute times)."""
import ctypes as C
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from manipulation_planning_suite_b200 import api
sc, start, batches, goals = bench.make_workload([0], 3)
names = ["broad phase", "narrow phase", "round scheduling", "solver rounds", "friction+schedule+solve+integrate", "contact steps"]
for bi in range(int(os.environ.get('PHASE_BATCHES', '3'))):
    c, n, d = batches[bi]
    ctx = api.Context(sc)
    ctx.extend_upload(start, c, n_ctrl=n, dest=d, compare_f32=True, goals=goals)
    for _ in range(3):
        ctx.extend_launch()
    r = ctx.extend_fetch(dump=True)
    K = c.shape[0]
    out = np.zeros(K * 12, np.int64)
    ctx._check(ctx.lib.mps_extend_chain_cycles(ctx.h, out.ctypes.data_as(C.POINTER(C.c_int64))))
    out = out.reshape(K, 12)
    steps = r.all_steps.sum(axis=1)
    order = np.argsort(-out[:, 0])[:6]
    print(f"batch {bi}: kernel-critical chains")
    for k in order:
        tot = out[k, 0]
        ph = out[k, 1:7]
        other = ph[4] - ph[2] - ph[3]
        print(f"  chain {k}: {tot} cyc ({tot/1.965e3:.0f} us) {steps[k]} steps ({ph[5]} with contacts), cyc/step {tot/max(steps[k],1):.0f} | "
              f"broad {ph[0]/tot:.1%} narrow {ph[1]/tot:.1%} sched {ph[2]/tot:.1%} solve {ph[3]/tot:.1%} "
              f"friction+integrate {other/tot:.1%} outside step {1-(ph[0]+ph[1]+ph[4])/tot:.1%}"
              f" | per contact step: solve {ph[3]/max(ph[5],1):.0f} ({out[k,10]/max(ph[5],1):.1f} passes of {ph[3]/max(out[k,10],1):.0f} cyc) sched {ph[2]/max(ph[5],1):.0f}; per step: broad {ph[0]/steps[k]:.0f} narrow {ph[1]/steps[k]:.0f} = reach {out[k,7]/steps[k]:.0f} + collide {out[k,8]/steps[k]:.0f} + rows {out[k,9]/steps[k]:.0f}; loops of one-round sweeps {out[k,11]} cyc")
    # the cheapest steps: chains that (almost) never touch anything
    free = [k for k in range(K) if steps[k] >= 100 and out[k, 6] <= 0.02 * steps[k]]
    free.sort(key=lambda k: out[k, 0] / steps[k])
    for k in free[:4] + free[-2:]:
        tot = out[k, 0]
        ph = out[k, 1:7]
        print(f"  free chain {k}: {steps[k]} steps ({ph[5]} with contacts), cyc/step {tot/steps[k]:.0f}: broad {ph[0]/steps[k]:.0f} "
              f"narrow {ph[1]/steps[k]:.0f} (reach {out[k,7]/steps[k]:.0f} collide {out[k,8]/steps[k]:.0f} rows {out[k,9]/steps[k]:.0f}) "
              f"friction+integrate {(ph[4]-ph[2]-ph[3])/steps[k]:.0f} outside the step {(tot-(ph[0]+ph[1]+ph[4]))/steps[k]:.0f}")
    tot = out[:, 0].sum()
    ph = out[:, 1:7].sum(axis=0)
    print(f"  all chains: broad {ph[0]/tot:.1%} narrow {ph[1]/tot:.1%} sched {ph[2]/tot:.1%} solve {ph[3]/tot:.1%} "
          f"friction+integrate {(ph[4]-ph[2]-ph[3])/tot:.1%} outside {1-(ph[0]+ph[1]+ph[4])/tot:.1%}; mean chain {out[:,0].mean()/1.965e3:.0f} us")
    ctx.close()
