"""Three serial post-process calls of C3 (LVIS-1203, A=6, b32): per-kernel launch list of the LVIS configuration."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from object_detectors_b200 import ops, synthetic as syn
base = [torch.from_numpy(h).cuda() for h in syn.yolo_heads(3001, 8, 608, 1203, syn.LVIS_ANCHORS, "clustered")]
heads = [torch.cat([h] * 4, 0).contiguous() for h in base]
idf = torch.from_numpy(np.load("tests/golden/idf_lvis_smooth.npy")).cuda()
plan = ops.YoloPostprocess([19, 38, 76], 32, syn.LVIS_ANCHORS, 608, 1203, True, 0.1, 0.6, ops.NMS_MAJORITY, 4096, 512, "cuda")
for _ in range(3):
    plan(heads, idf)
torch.cuda.synchronize()
plan.check_status()
print("ok", int(plan.cand_count.sum()), int(plan.cand_count.max()), int(plan.det_count.sum()))
