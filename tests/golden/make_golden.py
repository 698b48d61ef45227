"""Generate the frozen golden vectors in this directory from the REAL reference.

Run in the build container only (needs the read-only mount):

    python tests/golden/make_golden.py

It imports ``libcity.model.traffic_flow_prediction.MultiATGCN`` from
``/root/reference`` unmodified, builds it on small seeded synthetic graphs,
runs ``calculate_loss`` forward + backward on CPU in eval mode (dropout off, so
the numbers are deterministic) and stores inputs, the full ``state_dict``, the
forecast, the loss and every parameter gradient in one ``.npz`` per case.  The
GPU box has no ``/root/reference``; tests there read only these files.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from multistgraph_b200.synthetic import make_batch, make_config, make_data_feature  # noqa: E402

# name -> overrides.  N, B and widths are kept small so the files stay a few hundred KB.
CASES = {
    "multi_bi_k2": dict(N=11, B=2, cfg=dict(adjtype="multi", adpadj="bidirection", cheb_order=2, output_window=6)),
    "od_bi_k2": dict(N=9, B=3, cfg=dict(adjtype="od", adpadj="bidirection", cheb_order=2, output_window=3)),
    "multi_none_k2": dict(N=10, B=2, cfg=dict(adjtype="multi", adpadj="none", cheb_order=2, output_window=6)),
    "od_uni_k3": dict(N=8, B=2, cfg=dict(adjtype="od", adpadj="unidirection", cheb_order=3, output_window=6)),
    "multi_bi_k3": dict(N=8, B=2, cfg=dict(adjtype="multi", adpadj="bidirection", cheb_order=3, output_window=12)),
    "cosine_none_k1": dict(N=7, B=2, cfg=dict(adjtype="cosine", adpadj="none", cheb_order=1, output_window=6)),
    "dist_bi_k2": dict(N=9, B=2, cfg=dict(adjtype="dist", adpadj="bidirection", cheb_order=2, output_window=6)),
    "identity_bi_k2": dict(N=6, B=2, cfg=dict(adjtype="identity", adpadj="bidirection", cheb_order=2,
                                               output_window=6)),
    "gcn_off": dict(N=7, B=2, cfg=dict(adjtype="multi", adpadj="bidirection", cheb_order=2, output_window=6,
                                        gcn_off=True)),
    "fnn_off": dict(N=7, B=2, cfg=dict(adjtype="od", adpadj="bidirection", cheb_order=2, output_window=6,
                                        fnn_off=True)),
    "node_specific_off": dict(N=7, B=2, cfg=dict(adjtype="multi", adpadj="bidirection", cheb_order=2,
                                                  output_window=6, node_specific_off=True)),
    "one_layer": dict(N=7, B=2, cfg=dict(adjtype="multi", adpadj="bidirection", cheb_order=2, output_window=6,
                                          num_layers=1)),
    # add_static=true: static node features -> similarity view (MA.py:244-250), PCA-initialised node embedding
    # (MA.py:286-294) and initial hidden state (MA.py:335-338, 405-409); torch.pca_lowrank patched deterministic
    "add_static": dict(N=9, B=2, static_dim=6, cfg=dict(adjtype="multi", adpadj="bidirection", cheb_order=2,
                                                        output_window=6)),
}


def build_case(name, spec, seed=1234):
    sys.path.insert(0, "/root/reference")
    from libcity.model.traffic_flow_prediction.MultiATGCN import MultiATGCN as RefModel

    from tests.util import exact_pca_lowrank

    torch.pca_lowrank = exact_pca_lowrank  # the reference's randomised PCA, made deterministic (see tests/util.py)
    cfg = make_config(embed_dim=4, rnn_units=8, batch_size=spec["B"], **spec["cfg"])
    df = make_data_feature(spec["N"], seed=seed)
    if spec.get("static_dim"):
        df["static"] = np.random.default_rng(seed + 1).normal(size=(spec["N"], spec["static_dim"])).astype(np.float32)
    batch = make_batch(spec["N"], spec["B"], cfg["output_window"], seed=seed)
    torch.manual_seed(seed)
    model = RefModel(dict(cfg), df)
    model.eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    y = model.predict({"X": batch["X"].clone(), "y": batch["y"].clone()})
    loss = model.calculate_loss({"X": batch["X"].clone(), "y": batch["y"].clone()})
    loss.backward()
    out = {"X": batch["X"].numpy(), "y": batch["y"].numpy(), "adj_mx": df["adj_mx"],
           "geo_id": np.asarray(df["coordinate"]["geo_id"]),
           "coordinates": np.asarray(list(df["coordinate"]["coordinates"])).astype("U"),
           "forecast": y.detach().numpy(), "loss": np.asarray(loss.item(), dtype=np.float64)}
    if df.get("static") is not None:
        out["static"] = df["static"]
    for k, v in sd.items():
        out["param/" + k] = v.numpy()
    for k, p in model.named_parameters():
        if p.grad is not None:
            out["grad/" + k] = p.grad.numpy()
    for i, pair in enumerate(model.supports):
        out["support/%d" % i] = pair[1].numpy()
    cfg_json = {k: (str(v) if isinstance(v, torch.device) else v) for k, v in cfg.items()}
    out["config_json"] = np.asarray(json.dumps(cfg_json))
    out["len_windows"] = np.asarray([df["len_closeness"], df["len_period"], df["len_trend"]])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "forecast", y.shape, "loss %.6f" % loss.item(), "params", len(sd))


if __name__ == "__main__":
    only = sys.argv[1:]
    for nm, sp in CASES.items():
        if not only or nm in only:
            build_case(nm, sp)
