"""Generate the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports ``my_models.lpsr.LPSR`` from /root/reference (read-only), loads the shipped
``weights/best_model.pth`` (sha256 71d215d0...bb7dc8), runs CPU fp32 forwards on seeded inputs and
stores input/output pairs (+ per-stage intermediates captured with forward hooks) as small
compressed ``.npz`` files.  The oracle (oracle/lpsr_oracle.py), the torch port and the CUDA path
are all checked against these files; nothing reads /root/reference at test time.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from my_models.lpsr import LPSR  # noqa: E402  (the reference itself)

from oracle import lpsr_oracle as orc  # noqa: E402
from oracle import lpsr_torch_port as port  # noqa: E402

torch.set_num_threads(8)
torch.manual_seed(0)


def ref_model():
    return LPSR(3, 32, 16, 4, 4, None).eval()


def full_state_dict_from_live(model, live):
    """Expand 64 live tensors to the reference's 158-key layout (aliases + dead RDBs keep their init)."""
    sd = model.state_dict()
    out = {}
    for k, v in sd.items():
        kk = k
        if k.startswith("auto_encoder.GA.0."):
            kk = "auto_encoder.encoder." + k[len("auto_encoder.GA.0."):]
        elif k.startswith("auto_encoder.GA.1."):
            kk = "auto_encoder.decoder." + k[len("auto_encoder.GA.1."):]
        elif k.startswith("rdn.rdbs."):
            idx = int(k.split(".")[2])
            if idx % 2 == 1:
                kk = "rdn.csar." + ".".join(k.split(".")[3:])
        out[k] = torch.as_tensor(np.asarray(live[kk])).reshape(v.shape) if kk in live else v.clone()
    return out


def smooth_input(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    lo = torch.rand(b, 3, max(h // 8, 1), max(w // 8, 1), generator=g)
    x = torch.nn.functional.interpolate(lo, size=(h, w), mode="bicubic", align_corners=False)
    return x.clamp(0, 1).contiguous()


def uniform_input(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, 3, h, w, generator=g)


def capture_taps(model, x):
    taps = {}
    hooks = []

    def save(name, multi=False):
        def fn(_m, _i, o):
            if multi:
                j = sum(1 for k in taps if k.startswith(name + "#"))
                taps[f"{name}#{j}"] = o.detach().clone().numpy()
            else:
                taps[name] = o.detach().clone().numpy()
        return fn

    ae = model.auto_encoder
    hooks.append(ae.conv_in.register_forward_hook(save("ae.c0")))
    hooks.append(ae.encoder[2].register_forward_hook(save("ae.enc0")))
    hooks.append(ae.encoder[5].register_forward_hook(save("ae.enc1")))
    hooks.append(ae.decoder[2].register_forward_hook(save("ae.dec0")))
    hooks.append(ae.decoder[5].register_forward_hook(save("ae.dec1")))
    hooks.append(ae.register_forward_hook(save("ae.out")))
    hooks.append(model.rdn.shallowF1.register_forward_hook(save("rdn.sfe1")))
    hooks.append(model.rdn.shallowF2.register_forward_hook(save("rdn.sfe2")))
    hooks.append(model.rdn.rdbs[0].register_forward_hook(save("rdn.block0")))
    hooks.append(model.rdn.rdbs[2].register_forward_hook(save("rdn.block2")))
    hooks.append(model.rdn.csar.register_forward_hook(save("rdn.csar_out", multi=True)))
    hooks.append(model.rdn.csar.conv_in.register_forward_hook(save("rdn.csar_x_in", multi=True)))
    hooks.append(model.rdn.register_forward_hook(save("rdn.out")))
    with torch.no_grad():
        y = model(x)
    for h in hooks:
        h.remove()
    ren = {"rdn.csar_out#0": "rdn.block1", "rdn.csar_out#1": "rdn.block3",
           "rdn.csar_x_in#0": "csar1.x_in", "rdn.csar_x_in#1": "csar3.x_in"}
    taps = {ren.get(k, k): v for k, v in taps.items()}
    return y.numpy(), taps


def main():
    ck = os.path.join(REF, "weights", "best_model.pth")
    sha = hashlib.sha256(open(ck, "rb").read()).hexdigest()
    sd = torch.load(ck, map_location="cpu")
    sd = sd.get("model_state_dict", sd)
    model = ref_model()
    print(model.load_state_dict(sd, strict=True))

    live_names = orc.live_weight_names()
    live = {k: sd[k].numpy().astype(np.float32) for k in live_names}
    np.savez_compressed(os.path.join(HERE, "weights_best_model.npz"), **live)

    # 158-key layout: names, shapes, alias groups (by bit-identity in the checkpoint)
    keys = [{"name": k, "shape": list(v.shape)} for k, v in sd.items()]
    meta = {"checkpoint_sha256": sha, "n_keys": len(sd), "keys": keys,
            "torch": torch.__version__, "live_names": live_names}
    json.dump(meta, open(os.path.join(HERE, "state_dict_layout.json"), "w"), indent=0)

    Wt = port.to_torch_weights(live)
    cases = {
        "u_b2_32x192": uniform_input(2, 32, 192, 1234),       # call-site native crop (run.py:81)
        "u_b1_64x192": uniform_input(1, 64, 192, 0),          # BASELINE configs 1-3 crop
        "s_b1_64x192": smooth_input(1, 64, 192, 7),
        "u_b1_30x190_pad": uniform_input(1, 30, 190, 5),      # pad-to-4 path (lpsr.py:107-111)
        "u_b1_33x193_pad": uniform_input(1, 33, 193, 6),
        "u_b2_4x4": uniform_input(2, 4, 4, 3),
        "u_b1_128x384": uniform_input(1, 128, 384, 11)[:, :, :, :],  # config 4 crop
        # smooth crops are what real plates look like after the bicubic resize of the call site, and the class on which 16-bit
        # operand rounding is amplified most: a batch at the benchmark crop, the call-site crop and the pad-to-4 path
        "s_b8_64x192": smooth_input(8, 64, 192, 21),
        "s_b4_32x192": smooth_input(4, 32, 192, 22),
        "s_b2_30x190_pad": smooth_input(2, 30, 190, 23),
    }
    # zeros / ones / impulse in one batch (pins padding + shuffle indexing)
    sp = torch.zeros(3, 3, 16, 24)
    sp[1] = 1.0
    sp[2, 1, 5, 7] = 1.0
    cases["special_b3_16x24"] = sp
    for name, x in cases.items():
        with torch.no_grad():
            y = model(x)
        yp = port.lpsr_forward(x, Wt)
        assert torch.equal(y, yp), f"torch port differs from reference on {name}: {(y - yp).abs().max()}"
        yo = orc.lpsr_forward(x.numpy(), live)
        err = float(np.abs(yo - y.numpy()).max())
        print(f"{name:>20s}: out {tuple(y.shape)}  sum {float(y.sum()):.4f}  numpy-oracle max|d| {err:.2e}")
        assert err < 2e-5
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), x=x.numpy(), y=y.numpy())

    # per-stage intermediates on a small crop, shipped weights
    x = uniform_input(1, 16, 32, 21)
    y, taps = capture_taps(model, x)
    np.savez_compressed(os.path.join(HERE, "taps_u_b1_16x32.npz"), x=x.numpy(), y=y, **taps)
    print("taps:", sorted(taps))

    # random-init weights + adversarial alpha: weights regenerated from oracle.random_weights(seed)
    for tag, seed, alphas in (("rand_seed3", 3, None), ("rand_seed4_alpha", 4, (-2.5, 3.0))):
        lw = orc.random_weights(seed)
        if alphas is not None:
            lw["rdn.rdbs.0.alpha"] = np.float32(alphas[0]).reshape(())
            lw["rdn.rdbs.2.alpha"] = np.float32(alphas[1]).reshape(())
        m2 = ref_model()
        m2.load_state_dict(full_state_dict_from_live(m2, lw), strict=True)
        x = uniform_input(2, 32, 64, 100 + seed)
        with torch.no_grad():
            y = m2(x)
        yo = orc.lpsr_forward(x.numpy(), lw)
        print(f"{tag}: numpy-oracle max|d| {np.abs(yo - y.numpy()).max():.2e}")
        np.savez_compressed(os.path.join(HERE, f"case_{tag}_b2_32x64.npz"), x=x.numpy(), y=y.numpy(),
                            seed=np.int64(seed), alphas=np.asarray(alphas if alphas else [], dtype=np.float32))

    # the survey's sanity vector
    x = uniform_input(2, 64, 192, 1234)
    with torch.no_grad():
        y = model(x)
    print("sanity: sum", float(y.sum()), "mean", float(y.mean()), y[0, 0, 0, :4].tolist(), y[1, 0, 63, 188:].tolist())


if __name__ == "__main__":
    main()
