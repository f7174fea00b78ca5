"""TEST HELPER (not product code): runs an UNMODIFIED reference script (evaluate_ood.py) through
rba_b200.compat.plug_in().  With a GPU the real rba_b200 engine serves the model; without one (the build container)
the model's forward is replaced by the ORACLE's CPU forward so that everything around the hot path — config,
registries, checkpoint loading, datasets, OODEvaluator, metrics, results.pkl — is exercised end to end.

    python tests/helpers/run_reference_script.py <reference_root> <script.py> [script args...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402

from rba_b200 import compat  # noqa: E402

ref_root, script = sys.argv[1], sys.argv[2]
served = compat.plug_in()
print("[helper] stand-ins serve:", served)
if not torch.cuda.is_available():
    import rba_b200
    import rba_oracle as O

    def oracle_forward(self, batched_inputs, **kwargs):
        imgs = [x["image"].cpu() for x in batched_inputs]
        ref = O.forward(self.state_dict(), self.mc, imgs)
        return [{"sem_seg": s} for s in ref["sem_seg"]]

    rba_b200.MaskFormer.forward = oracle_forward
    print("[helper] no GPU: rba_b200.MaskFormer.forward -> oracle CPU forward (test only)")
os.chdir(ref_root)
from rba_b200.compat.run import run_script  # noqa: E402

run_script(os.path.join(ref_root, script), sys.argv[3:])
