"""The class LibCity's model registry should find.

``get_model`` (libcity/utils/utils.py:46-50) does ``getattr(importlib.import_module('libcity.model.traffic_flow_prediction'),
config['model'])(config, data_feature)``, and that package re-exports ``MultiATGCN`` from MultiATGCN.py
(libcity/model/traffic_flow_prediction/__init__.py:1-6).  The whole reference-side change is that one import line::

    # libcity/model/traffic_flow_prediction/__init__.py
    from multistgraph_b200.libcity_plugin import MultiATGCN

This module needs ``libcity`` on the path (it is imported from inside the LibCity tree) and gives the accelerated model the
base class the reference's own model has (``AbstractTrafficStateModel``, libcity/model/abstract_traffic_state_model.py:4-30;
MultiATGCN.py:221), so ``isinstance`` checks and anything the harness adds to that base keep working.  Everything else
(ConfigParser, MTHDataset, TrafficStateExecutor, TrafficStateEvaluator, run_model.py) runs unmodified:
tests/test_dropin_run_model.py drives ``run_model`` through a scratch mirror of the reference tree with this line in place.
"""
from libcity.model.abstract_traffic_state_model import AbstractTrafficStateModel

from multistgraph_b200.model import MultiATGCN as _Accelerated


class MultiATGCN(_Accelerated, AbstractTrafficStateModel):
    def __init__(self, config, data_feature):
        _Accelerated.__init__(self, config, data_feature)


__all__ = ["MultiATGCN"]
