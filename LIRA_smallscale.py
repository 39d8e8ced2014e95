#!/usr/bin/env python
"""LIRA for small-scale datasets on liblira_b200 -- same entry point, argv and module-level names as the reference's
LIRA_smallscale.py (Config :27-75, mul_partition_by_model :77-97, cal_metrics :99-143, get_cmp_recall :145-174,
query_tuning :176-241, `__main__` :246-379):

    python LIRA_smallscale.py --dataset sift --n_bkt 1024 --k 10 [--dis_metric L2] [--data_path /data/vector_datasets]

The functions live in the package (lira-ann-search_b200/drivers.py, query.py, utils.py); this module re-exports them so that
`import LIRA_smallscale as S; S.get_cmp_recall(...)` keeps working.
"""
import lira_ann_search_b200  # noqa: F401  (registers the package under its importable name)
from lira_ann_search_b200.drivers import Config, cal_metrics, parse_config, run_smallscale
from lira_ann_search_b200.model_probing import MLP_2_Input, model_evaluate, model_infer, model_train  # noqa: F401
from lira_ann_search_b200.query import get_cmp_recall, mul_partition_by_model, query_tuning  # noqa: F401
from lira_ann_search_b200.utils import *  # noqa: F401,F403

if __name__ == "__main__":
    run_smallscale(parse_config(Config))
