"""Timeline of the fused logistic sweep's CTA 0 (RMN_LGF_TIMELINE=1 RMN_LGF_TIMELINE_FILE=...): per tile, clock64 stamps of
the MMA thread (after the ring-A wait, GEMM1 issued, R of the tile ready, GEMM2 issued) and of one pointwise warp (logits
ready, tcgen05.ld done, arithmetic done, R handed over).  Prints per-phase medians in cycles.
    python scripts/lgf_timeline.py gpurun_out/tl.bin"""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.int64).reshape(256, 8)
ok = np.all(t[:, :8] > 0, axis=1)
t = t[ok][8:200]
names = ["g1_after_fullA", "g1_issued", "rfull_ready", "g2_issued", "pw_zfull", "pw_ld_done", "pw_math_done", "pw_arrived"]
per = np.median(np.diff(t[:, 1]))
print("tiles with stamps:", len(t), " median cycles per tile:", per)
def med(a, b, la):
    print("%-46s %8.0f" % (la, np.median(t[:, b] - t[:, a])))
med(0, 1, "GEMM1 issue (27 MMAs + commit), same tile")
med(1, 4, "GEMM1 issued -> logits ready at the pw warp")
med(4, 5, "tcgen05.ld + wait")
med(5, 6, "pointwise arithmetic (8 pairs)")
med(6, 7, "tcgen05.st + wait + fence + arrive")
med(7, 2, "pw arrived -> GEMM2 warp sees R (all 16 warps)")
med(2, 3, "GEMM2 issue (4 MMAs + 2 commits, own warp)")
print("%-46s %8.0f" % ("GEMM2(t) issued -> GEMM1(t+2) starts (z_free, fullA)", np.median(t[2:, 0] - t[:-2, 3])))
print("%-46s %8.0f" % ("GEMM1(t+1) issued -> R(t) seen by the GEMM2 warp", np.median(t[:-1, 2] - t[1:, 1])))
