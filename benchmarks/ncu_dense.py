"""Three dense `[B,N,85]` decodes of the C2 batch: the target of the ncu captures of k_decode_dense2."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from object_detectors_b200 import ops, synthetic as syn
heads = [torch.from_numpy(h).cuda() for h in syn.yolo_heads(1000, 64, 608, 80, syn.COCO_ANCHORS, "clustered")]
idf = torch.from_numpy(np.load("tests/golden/idf_coco_smooth.npy")).cuda()
for _ in range(3):
    ops.yolo_decode_dense(heads, syn.COCO_ANCHORS, 608, 80, idf, True)
torch.cuda.synchronize()
print("ok")
