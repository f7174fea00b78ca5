"""CPU, build container only: oracle vs the LIVE reference (skipped where /root/reference is absent)."""
import pytest
import torch
import yaml

import rba_oracle as O
import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")


@pytest.mark.parametrize("name", ["swin_b_1dl"])
def test_live_reference(name):
    cfgr = ref_loader.load_cfg(name)
    m = ref_loader.build_reference_model(cfgr)
    sd = O.perturb_state_dict(m.state_dict())
    m.load_state_dict(sd)
    cfg = O.config_from_yaml_dict(yaml.safe_load(open(f"{ref_loader.REF_ROOT}/ckpts/{name}/config.yaml")))
    g = torch.Generator().manual_seed(1)
    imgs = [torch.randint(0, 256, (3, 72, 100), dtype=torch.uint8, generator=g) for _ in range(2)]
    with torch.no_grad():
        ref = m([{"image": im} for im in imgs])
        mine = O.forward(sd, cfg, imgs)
    for b in range(2):
        assert (ref[b]["sem_seg"] - mine["sem_seg"][b]).abs().max() < 5e-5
        assert (-ref[b]["sem_seg"].tanh().sum(0) - mine["rba"][b]).abs().max() < 5e-5
