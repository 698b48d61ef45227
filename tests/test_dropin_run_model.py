"""The accelerated model as a drop-in under the reference's own ``run_model`` (BASELINE north star; run_model.py:8-29 ->
pipeline.py:16-62 -> utils.py:46-50): unmodified ConfigParser, MTHDataset, TrafficStateExecutor (train with early stopping,
checkpoints, LR schedule; evaluate with the group-std re-transform) and TrafficStateEvaluator drive
``multistgraph_b200.libcity_plugin.MultiATGCN`` on a synthetic dataset in LibCity's atomic-file format.

Runs where the reference tree is mounted (this container, CPU): the three C-ABI operators are swapped for the torch mirror
of the kernel algorithm (tests/host_mirror.py), exactly as tests/test_model_host.py does - the product path itself has no CPU
fallback.  (``pad_with_last_sample`` is switched off in the config: the reference's padding code, data/utils.py:55, builds a ragged
numpy array, which numpy >= 1.24 rejects; R2 / EVAR are left out of the evaluator's metric list: loss.py's sklearn wrappers return a
float under the installed scikit-learn and traffic_state_evaluator.py:116 calls .item() on it.)  On the GPU box the reference is absent (nothing there may read /root/reference); the same harness calls are
covered there by tests/test_gpu_train.py::test_fused_train_step_matches_executor_loop_on_the_model."""
import glob
import importlib
import os
import sys

import pytest
import torch

from tests import dropin, host_mirror

pytestmark = pytest.mark.skipif(not dropin.reference_available(), reason="reference tree not mounted")
DATASET = "SYN_DC_SHAPE_HOURLY"


@pytest.fixture()
def scratch(tmp_path, monkeypatch):
    root = str(tmp_path)
    dropin.make_scratch_tree(root)
    dropin.write_dataset(root, DATASET)
    dropin.install_stubs()
    monkeypatch.chdir(root)
    monkeypatch.syspath_prepend(root)
    stale = [m for m in sys.modules if m == "libcity" or m.startswith("libcity.") or m == "multistgraph_b200.libcity_plugin"]
    for m in stale:
        monkeypatch.delitem(sys.modules, m)
    yield root
    for m in [m for m in sys.modules if m == "libcity" or m.startswith("libcity.") or m == "multistgraph_b200.libcity_plugin"]:
        sys.modules.pop(m, None)


def test_registry_returns_the_accelerated_class(scratch):
    from libcity.config import ConfigParser
    from libcity.data import get_dataset
    from libcity.model.abstract_traffic_state_model import AbstractTrafficStateModel
    from libcity.utils import get_model

    from multistgraph_b200.model import MultiATGCN as Accelerated

    config = ConfigParser("traffic_state_pred", "MultiATGCN", DATASET, "config_user", False, True,
                          {"gpu": False, "batch_size": 4, "output_window": 3, "adjtype": "multi", "adpadj": "bidirection",
                           "embed_dim_node": 4, "embed_dim_adj": 4, "rnn_units": 8, "exp_id": 1, "pad_with_last_sample": False})
    dataset = get_dataset(config)
    dataset.get_data()
    model = get_model(config, dataset.get_data_feature())
    assert isinstance(model, Accelerated) and isinstance(model, AbstractTrafficStateModel)
    assert type(model).__module__ == "multistgraph_b200.libcity_plugin"
    ref = importlib.import_module("libcity.model.traffic_flow_prediction.MultiATGCN").MultiATGCN(config, dataset.get_data_feature())
    assert [k for k, _ in model.named_parameters()] == [k for k, _ in ref.named_parameters()]
    assert [tuple(p.shape) for p in model.parameters()] == [tuple(p.shape) for p in ref.parameters()]


def test_run_model_trains_and_evaluates_through_the_unmodified_executor(scratch):
    from libcity.pipeline import run_model

    from multistgraph_b200 import ops

    restore = host_mirror.install(ops)   # CPU stand-in for the three CUDA operators (tests only)
    try:
        run_model(task="traffic_state_pred", model_name="MultiATGCN", dataset_name=DATASET, config_file="config_user",
                  saved_model=True, train=True,
                  other_args={"gpu": False, "batch_size": 4, "output_window": 3, "max_epoch": 2, "adjtype": "multi",
                              "adpadj": "bidirection", "embed_dim_node": 4, "embed_dim_adj": 4, "rnn_units": 8, "exp_id": 7,
                              "seed": 0, "pad_with_last_sample": False,
                              "metrics": ["MAE", "MAPE", "MSE", "RMSE", "masked_MAE", "masked_MAPE", "masked_MSE", "masked_RMSE"]})
    finally:
        restore()
    out = os.path.join(scratch, "libcity", "cache", "7")
    assert glob.glob(os.path.join(out, "model_cache", "*.m")), "executor.save_model wrote no checkpoint"
    csvs = [f for f in glob.glob(os.path.join(out, "evaluate_cache", "*.csv")) if not f.endswith("_trans.csv")]
    assert csvs, "evaluator wrote no metric table"
    import pandas as pd
    table = pd.read_csv(csvs[0])
    assert len(table) == 3 and table["MAE"].notna().all() and (table["MAE"] > 0).all()
    state, opt_state = torch.load(glob.glob(os.path.join(out, "model_cache", "*.m"))[0], weights_only=False)
    assert "encoder.agru_cells.0.gate.weights_pool" in state and "node_emb" in state
