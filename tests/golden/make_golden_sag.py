#!/usr/bin/env python
"""Golden fixtures for the SAG decoder by RUNNING THE REFERENCE module
(/root/reference/scripts/model/motionclip_module.py::Decoder_TRANSFORMER, torch's own nn.TransformerDecoder) on CPU,
next to the oracle's restatement (oracle/sag_oracle.py).  The reference forward hard-codes `.cuda()` for one
temporary (motionclip_module.py:161); it is mapped to a no-op for this run.  Writes tests/golden/sag.npz.

    python tests/golden/make_golden_sag.py
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from livelyspeaker_b200 import synthetic     # noqa: E402
from oracle import sag_oracle                 # noqa: E402


def main():
    sys.path.insert(0, "/root/reference/scripts")
    sys.modules["clip"] = types.ModuleType("clip")
    from model.motionclip_module import Decoder_TRANSFORMER
    torch.manual_seed(7)
    ref = Decoder_TRANSFORMER(latent_dim=512, n_pre_poses=4, use_style=False).eval()
    sd = synthetic.synth_sag_state_dict(seed=3)
    missing, unexpected = ref.load_state_dict(sd, strict=True), None
    B = 3
    g = torch.Generator().manual_seed(21)
    x = 0.3 * torch.randn(B, 9, 3, 34, generator=g)
    z = torch.randn(B, 512, generator=g)
    z = z / z.norm(dim=-1, keepdim=True) * 10.0          # CLIP-feature-like scale
    mask = torch.ones(B, 34, dtype=torch.bool)
    mask[1, 30:] = False                                    # padded tail on one clip
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with torch.no_grad():
            r = ref({"x": x.clone(), "z": z.clone(), "mask": mask})["output"]
    finally:
        torch.Tensor.cuda = orig_cuda
    with torch.no_grad():
        o = sag_oracle.decode(sd, x, z, mask)
    err = float((r - o).abs().max())
    print("max |reference - oracle| =", err, " max |out| =", float(r.abs().max()))
    assert err < 2e-5, err
    np.savez_compressed(os.path.join(HERE, "sag.npz"), x=x.numpy(), z=z.numpy(), mask=mask.numpy(), output=r.numpy(),
                        weights_abs_sum=np.float64(sum(float(v.double().abs().sum()) for v in sd.values())))
    rp = os.path.join(HERE, "PIN_REPORT.json")
    allrep = json.load(open(rp)) if os.path.exists(rp) else {}
    allrep["sag"] = {"decode": err, "keys": len(sd)}
    json.dump(allrep, open(rp, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
