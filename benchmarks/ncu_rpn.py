"""Three C5 RPN filter calls (b16, 800x1344, k=2000) and nothing else: the target of the ncu launch list /
`--set full` captures behind profiles/r02_rpn_kernels.csv and profiles/r02_nms_fused_notes.md.

    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_rpn|k_nms" python benchmarks/ncu_rpn.py
"""
import sys, os, torch
sys.path.insert(0, os.getcwd())
from object_detectors_b200 import ops, synthetic as syn
obj, deltas, anchors, per_level = syn.rpn_inputs(41, 16, 800, 1344)
to, td, ta = torch.from_numpy(obj).cuda(), torch.from_numpy(deltas).cuda(), torch.from_numpy(anchors).cuda()
hw = torch.tensor([[800, 1344]] * 16, dtype=torch.float32).cuda()
for _ in range(3):
    ops.rpn_filter(to, td, ta, per_level, hw, 2000, 2000, 0.7, 0.0, 1e-3, ops.NMS_TV_CLASS)
torch.cuda.synchronize()
print("ok")
