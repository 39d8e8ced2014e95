#!/usr/bin/env python
"""LIRA for large-scale datasets on liblira_b200 -- same entry point, argv and module-level names as the reference's
LIRA_largescale.py (Config :27-49, mul_partition_by_model :51-72, get_cmp_recall :120-149, query_tuning :151-179,
`__main__` :184-354): the probing model is trained on a 1 % subset, every point of the full data then gets its second
partition from the model in batches of 1 000 000 (full redundancy).

    python LIRA_largescale.py --dataset deep50M --n_bkt 1024 --k 100
"""
import lira_ann_search_b200  # noqa: F401
from lira_ann_search_b200.drivers import LargeConfig as Config
from lira_ann_search_b200.drivers import cal_metrics, parse_config, run_largescale  # noqa: F401
from lira_ann_search_b200.model_probing import MLP_2_Input, model_evaluate, model_infer, model_train  # noqa: F401
from lira_ann_search_b200.query import get_cmp_recall  # noqa: F401
from lira_ann_search_b200.query import mul_partition_by_model_large as mul_partition_by_model  # noqa: F401
from lira_ann_search_b200.query import query_tuning_large as query_tuning  # noqa: F401
from lira_ann_search_b200.utils import *  # noqa: F401,F403

if __name__ == "__main__":
    run_largescale(parse_config(Config))
