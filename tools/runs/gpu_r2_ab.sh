#!/bin/bash
# Round 2, call AB: the TMA-staged kernel with 8 compute warps (one CTA per SM).
python - <<'PY'
import torch, numpy as np
from ddm_b200 import _cabi, ops
import oracle
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 1e-2)):
    for regime in ("late", "early", "dups"):
        x0 = torch.randn(128, 3072, generator=g).clamp(-1, 1)
        xh = x0[:, None] + 0.05 * torch.randn(128, 8, 3072, generator=g) if regime == "late" else torch.randn(128, 8, 3072, generator=g)
        if regime == "dups":
            xh[:, 3] = xh[:, 1]
            xh[:, 5] = xh[:, 1] + 1e-4 * torch.randn(128, 3072, generator=g)
        xh, x0 = xh.to(dev).to(dtype), x0.to(dev).to(dtype)
        loss, conf, inter, grad = oracle.energy_loss(xh.double().cpu().numpy(), x0.double().cpu().numpy(), 0.1, 1.0, 0.5)
        w = torch.full((1,), 0.5 * 128, device=dev)
        for thr in (0, 256):
            _cabi.set_tuning("energy.threads", thr); _cabi.set_tuning("energy.nv", 1 if thr else 0)
            out, gr = ops.energy_fused(xh, x0, w, 1.0 / 128, 0.1, 1.0, True)
            err = np.max(np.abs(gr.float().cpu().numpy() - grad)) / np.max(np.abs(grad))
            print(str(dtype)[6:], regime, "threads", thr, _cabi.describe_energy(128, 8, 3072, "bf16" if dtype == torch.bfloat16 else "f32")[:60],
                  "loss_err %.2e grad_err %.2e" % (abs(float(out[0]) - loss) / abs(conf), err), "OK" if err <= tol else "FAIL")
_cabi.set_tuning("energy.threads", 0); _cabi.set_tuning("energy.nv", 0)
PY
echo "== f32 one stream"
timeout 300 python tools/sweep_energy.py --streams 1 --configs "variant=3;variant=3,threads=256,nv=1;variant=3,threads=256,nv=2;variant=3,threads=192,nv=1"
echo "== f32 six streams"
timeout 300 python tools/sweep_energy.py --streams 6 --configs "variant=3;variant=3,threads=256,nv=1"
echo "== bf16 one / six streams"
timeout 300 python tools/sweep_energy.py --streams 1 --dtype bf16 --configs "variant=3;variant=3,threads=256,nv=1"
timeout 300 python tools/sweep_energy.py --streams 6 --dtype bf16 --configs "variant=3;variant=3,threads=256,nv=1"
timeout 200 python tools/trace_energy.py --tune "energy.threads=256,energy.nv=1" | grep "us/launch\|->\|period" | head -12
