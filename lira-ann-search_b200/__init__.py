"""lira-ann-search_b200: B200-native query phase and ground-truth path of LIRA.

Host-side mirror of the reference's Python entry points over liblira_b200.so (hand-written sm_100a
CUDA behind a C ABI, include/lira_b200.h). Import as `lira_ann_search_b200` (see the shim module at
the repository root). No CPU fallback: compute calls raise when the library or a GPU is missing.
"""
from . import _cabi
from ._cabi import (METRIC_IP, METRIC_L2, SELECT_GE_ARGMAX, SELECT_GT, SELECT_TOPN, LiraError)
from .engine import KnnIndex, LiraIndex, LiraModel, ListView, centroid_features, knn, launch_count
from .model_probing import MLP_2_Input, model_evaluate, model_infer, model_train
from .query import (cpp_thresholds, get_cmp_recall, mul_partition_by_model, mul_partition_by_model_large, query_tuning,
                    query_tuning_large, recall_at_k, search_sweep)
from .drivers import Config, LargeConfig, cal_metrics, parse_config, run_largescale, run_smallscale
from .utils import (build_kmeans_index, compute_data_knn, create_flat_indexes, create_inner_indexes,
                    fprint, get_dist_cid, get_idle_gpu, get_knn_distr_redundancy, get_knn_labels_data_only,
                    get_scaled_dist, get_scaled_dist_data, load_data, per_query, read_xvecs, write_xvecs)

__all__ = [n for n in dir() if not n.startswith("_")]
