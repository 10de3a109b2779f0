"""sub-batch pipelines: enqueue-ahead (all steps queued in round-robin order) vs completion-ordered issue (a
sub-batch's next step is enqueued when its previous step has finished) (GPU)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
from mpc_implementation_b200.closed_loop import PipelinedClosedLoop
sc = b200nmpc.SCENARIOS['nmpc_tt']; B = 4096; K, W = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 3
for S in (4, 8):
    for mode in ("ahead", "completion", "depth2"):
        p, vw = b200nmpc.random_instances(sc, B, seed=2000)
        pl = PipelinedClosedLoop(lambda n: b200nmpc.nlpsol('s', 'ipm', sc, max_batch=n), sc, p, target_vw=vw, pipelines=S)
        for k in range(W): pl.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode == "ahead":
            for k in range(K): pl.step()
        else:
            depth = 1 if mode == "completion" else 2
            done = [0] * S; evs = [[] for _ in range(S)]
            def issue(i):
                with torch.cuda.stream(pl.streams[i]):
                    pl.loops[i].step(); e = torch.cuda.Event(); e.record(); evs[i].append(e)
                done[i] += 1
            for d in range(depth):
                for i in range(S): issue(i)
            while min(done) < K or any(evs):
                for i in range(S):
                    if evs[i] and evs[i][0].query():
                        evs[i].pop(0)
                        if done[i] < K: issue(i)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        print(f'S={S} {mode}: {ms/K:.2f} ms/step')
