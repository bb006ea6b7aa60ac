"""B2S_TRACE timelines of ICP.process_sequence / process_batch for different pipeline depths."""
import os, sys
os.environ["B2S_TRACE"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b2slam
from b2slam import synth, _lib


def pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


xy, _ = synth.room_sequence(9001, 10000, 360)
kq, hq = pinned(xy)
ka, ha = pinned(xy[:-1])
kb, hb = pinned(xy[1:])
icp = b2slam.ICP()
icp.process_sequence(hq)
for chunks in (1, 2, 4, 8):
    _lib.check(_lib.lib().b2s_tune(b"h2d_chunks", chunks))
    for rep in range(2):
        print("== process_sequence chunks %d rep %d" % (chunks, rep), file=sys.stderr, flush=True)
        icp.process_sequence(hq)
_lib.check(_lib.lib().b2s_tune(b"h2d_chunks", 0))
for rep in range(2):
    print("== process_batch rep %d" % rep, file=sys.stderr, flush=True)
    icp.process_batch(ha, hb)
