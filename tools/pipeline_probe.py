"""does splitting the batch into independently pipelined sub-batches (own handle + stream each) hide the straggler
tail of one sub-batch behind the bulk of the other? (GPU)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
from mpc_implementation_b200.closed_loop import ClosedLoop
sc = b200nmpc.SCENARIOS['nmpc_tt']; B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K, W = 20, 3
for S in (1, 2, 3, 4, 8):
    p, vw = b200nmpc.random_instances(sc, B, seed=2000)
    idx = np.array_split(np.arange(B), S)
    streams = [torch.cuda.Stream() for _ in range(S)]
    loops = []
    for i in range(S):
        with torch.cuda.stream(streams[i]):
            s = b200nmpc.nlpsol('s', 'ipm', sc, max_batch=len(idx[i]))
            loops.append(ClosedLoop(s, sc, p[idx[i]], target_vw=vw[idx[i]]))
    torch.cuda.synchronize()
    conv = 0
    for k in range(W + K):
        if k == W:
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            conv_t = [torch.zeros((), dtype=torch.int64, device='cuda') for _ in range(S)]
            e0.record()
            for st_ in streams: st_.wait_event(e0)
        for i in range(S):
            with torch.cuda.stream(streams[i]):
                loops[i].step(want_g=False, want_lam=False)
                if k >= W: conv_t[i] += loops[i].solver.stats()["success"].sum()
    for st_ in streams: torch.cuda.current_stream().wait_stream(st_)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    conv = sum(int(c.item()) for c in conv_t)
    print(f'B={B} sub-batches={S}: {ms/K:.2f} ms/step, {conv/(ms*1e-3):.0f} converged solves/s, converged {conv/(B*K):.4f}')
