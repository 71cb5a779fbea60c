"""Three ROI-head post-process calls (COCO-91 b16, LVIS-1204 b4, 1000 proposals each) for an ncu launch list."""
import sys, os, torch
sys.path.insert(0, os.getcwd())
from object_detectors_b200 import ops, synthetic as syn
for c, batch in ((91, 16), (1204, 4)):
    logits, regs, props = syn.roi_inputs(51, [1000] * batch, c, 800, 1216)
    lg, rg = torch.from_numpy(logits).cuda(), torch.from_numpy(regs).cuda()
    pr = [torch.from_numpy(p).cuda() for p in props]
    shapes = [(800, 1216)] * batch
    for _ in range(3):
        out = ops.roi_postprocess(lg, rg, pr, shapes, None, ops.ROI_SOFTMAX, capacity=8192)
    torch.cuda.synchronize()
    print("ok", c, int(out[3].max()), int(out[2].sum()))
