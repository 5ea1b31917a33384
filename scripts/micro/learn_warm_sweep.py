import sys, time
sys.path.insert(0, "detecting-convolutional-codes-via-markovian-statistics_b200")
import numpy as np
from mvd import bitsource
from mvd.engine import Detector, Seg
for name, gen in (("m4a", [[[1,1,0,0,1]],[[1,1,0,1,1]]]), ("m4c", [[[1,0,0,1,1]],[[1,1,1,0,1]]]), ("m3", [[[1,1,1,1]],[[1,0,1,1]]])):
    m = len(gen[0][0]) - 1
    det = Detector(gen, 1, 2, m, enumerate_with="gpu", max_states=1 << 20)
    L = 200 * det.S
    for p in (0.001, 0.1, 0.5):
        seg = Seg(N=L, threshold=bitsource.bsc_threshold(p), stream=bitsource.LEARN_STREAM, enc_taps=det.dec_taps)
        ref = None
        for warm in (128, 256, 384, 512):
            det.learn_warm(warm)
            det.learn_counts([seg], burn=200, seed=1)
            c = det.learn_counts([seg], burn=200, seed=1)
            if ref is None: ref = c
            assert np.array_equal(c, ref)
            print(name, "S", det.S, "L", L, "p", p, "warm", warm, "ms %.3f" % det.last_kernel_ms(), "dirty", det.learn_dirty_chunks(), "of", (L + 127) // 128, flush=True)
    det.close()
