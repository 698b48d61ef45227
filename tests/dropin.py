"""Scratch mirror of the reference tree for the drop-in test (tests/test_dropin_run_model.py): LibCity reads every config and
data file relative to the working directory and /root/reference is read-only, so the harness is copied to a temporary
directory at test time (never into the repository), the ONE binding line of multistgraph_b200/libcity_plugin.py is put in
place, and a small synthetic dataset in LibCity's atomic-file format is written next to it (the real
.geo/.rel/.dyna/.ext/.gbst files are not redistributable).  ray / hyperopt are only imported, never used, by the default
pipeline: empty stand-in modules are enough.  numpy >= 1.24 dropped ``np.float`` (traffic_state_datatset.py:284, 975)."""
import json
import os
import shutil
import sys
import types

import numpy as np

REFERENCE = "/root/reference"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE, "libcity"))


def install_stubs():
    if not hasattr(np, "float"):
        np.float = float
    names = ["ray", "ray.tune", "ray.tune.suggest", "ray.tune.suggest.hyperopt", "ray.tune.suggest.bayesopt",
             "ray.tune.suggest.basic_variant", "ray.tune.schedulers", "hyperopt", "hyperopt.pyll", "hyperopt.pyll.base"]
    for n in names:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    attrs = {"ray": ["tune"], "ray.tune": ["suggest", "schedulers"],
             "ray.tune.suggest": ["ConcurrencyLimiter", "hyperopt", "bayesopt", "basic_variant"],
             "ray.tune.suggest.hyperopt": ["HyperOptSearch"], "ray.tune.suggest.bayesopt": ["BayesOptSearch"],
             "ray.tune.suggest.basic_variant": ["BasicVariantGenerator"],
             "ray.tune.schedulers": ["FIFOScheduler", "ASHAScheduler", "MedianStoppingRule"],
             "hyperopt": ["hp", "fmin", "tpe", "atpe", "rand", "pyll"], "hyperopt.pyll": ["base"], "hyperopt.pyll.base": ["Apply"]}
    for mod, names_ in attrs.items():
        for a in names_:
            full = mod + "." + a
            if not hasattr(sys.modules[mod], a):
                setattr(sys.modules[mod], a, sys.modules.get(full, type(a, (), {})))


def write_dataset(root, name, n_nodes=10, hours=24 * 31 + 6, seed=0):
    """<root>/raw_data/<name>/: config.json (the DC dataset's schema), .geo, .rel (OD counts), .dyna (hourly visits), .ext, .gbst."""
    from multistgraph_b200.synthetic import make_series

    rng = np.random.default_rng(seed)
    d = os.path.join(root, "raw_data", name)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(REFERENCE, "raw_data", "201901010601_DC_SG_CTractFIPS_Hourly_Single_GP", "config.json")) as f:
        cfg = json.load(f)
    for k in ("data_files", "geo_file", "rel_file", "ext_file"):
        cfg["info"][k] = [name] if k == "data_files" else name
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    ids = [11001000100 + 7 * i for i in range(n_nodes)]
    with open(os.path.join(d, name + ".geo"), "w") as f:
        f.write("geo_id,type,coordinates\n")
        for i in ids:
            f.write('%d,Point,"[%.6f, %.6f]"\n' % (i, -77.0 + 0.2 * rng.random(), 38.8 + 0.2 * rng.random()))
    with open(os.path.join(d, name + ".rel"), "w") as f:
        f.write("rel_id,type,origin_id,destination_id,link_weight\n")
        r = 0
        for a in ids:
            for b in ids:
                f.write("%d,geo,%d,%d,%.3f\n" % (r, a, b, (50.0 if a == b else 0.0) + 20.0 * rng.random() + 0.5))
                r += 1
    series = make_series(n_nodes, hours, seed=seed)[..., 0].numpy()
    t0 = np.datetime64("2019-01-01T06:00:00")
    times = [str(t0 + np.timedelta64(h, "h")) + "Z" for h in range(hours)]
    with open(os.path.join(d, name + ".dyna"), "w") as f:
        f.write("dyna_id,type,time,entity_id,Visits\n")
        r = 0
        for j, i in enumerate(ids):          # grouped by entity, time ascending (traffic_state_datatset.py:244-287)
            for h in range(hours):
                f.write("%d,state,%s,%d,%.4f\n" % (r, times[h], i, series[h, j]))
                r += 1
    with open(os.path.join(d, name + ".ext"), "w") as f:
        f.write("ext_id,time,holiday,weekend,temp,rain,snow,New_cases\n")
        for h in range(hours):
            f.write("%d,%s,0,%d,%.2f,0,0,0\n" % (h, times[h], 1 if ((h + 6) // 24) % 7 >= 5 else 0, 10.0 + 5.0 * np.sin(h / 24.0)))
    with open(os.path.join(d, name + ".gbst"), "w") as f:
        f.write("geo_id,All_m,All_std\n")
        for j, i in enumerate(ids):
            f.write("%d,%.4f,%.4f\n" % (i, series[:, j].mean(), series[:, j].std()))
    return ids


def make_scratch_tree(root):
    """Copies the harness (libcity/, config_user.json) and applies the one-line binding."""
    shutil.copytree(os.path.join(REFERENCE, "libcity"), os.path.join(root, "libcity"),
                    ignore=shutil.ignore_patterns("temp", "cache", "log", "__pycache__"))
    shutil.copy(os.path.join(REFERENCE, "config_user.json"), os.path.join(root, "config_user.json"))
    with open(os.path.join(root, "libcity", "model", "traffic_flow_prediction", "__init__.py"), "w") as f:
        f.write("from multistgraph_b200.libcity_plugin import MultiATGCN\n\n__all__ = [\"MultiATGCN\"]\n")
    os.makedirs(os.path.join(root, "libcity", "log"), exist_ok=True)
    os.makedirs(os.path.join(root, "libcity", "cache", "dataset_cache"), exist_ok=True)
