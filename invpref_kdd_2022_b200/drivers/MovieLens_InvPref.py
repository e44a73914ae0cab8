"""MovieLens_InvPref: InvPrefImplicit with the hyper-parameters of the reference driver
(reference MovieLens_InvPref.py:17-67).  Run: ``python -m invpref_kdd_2022_b200.drivers.MovieLens_InvPref [--epochs N] [--synthetic]``."""
import sys

from . import _common

MODEL_CONFIG: dict = {'env_num': 2, 'factor_num': 40, 'reg_only_embed': True, 'reg_env_embed': True}

TRAIN_CONFIG: dict = {'batch_size': 65536,
 'epochs': 4000,
 'cluster_interval': 20,
 'evaluate_interval': 10,
 'lr': 0.01,
 'invariant_coe': 8.909348155983732,
 'env_aware_coe': 1.233057369609993,
 'env_coe': 8.064376793624795,
 'L2_coe': 3.4987474005653665,
 'L1_coe': 0.9355983539586914,
 'alpha': None,
 'use_class_re_weight': False,
 'use_recommend_re_weight': True,
 'test_begin_epoch': 0,
 'begin_cluster_epoch': None,
 'stop_cluster_epoch': None}

EVALUATE_CONFIG: dict = {'top_k_list': [10, 20, 30], 'test_batch_size': 2048, 'eval_k': 30, 'eval_metric': 'ndcg'}

RANDOM_SEED_LIST = [17373331, 17373511, 17373423]

DATASET_PATH = '/MovieLens_all_data_thr_3/'
METRIC_LIST = ['ndcg', 'recall', 'precision']
USE_ITEM_POOL = False            # MovieLens_InvPref.py:90, :204: evaluator ranks all items
HAS_ITEM_POOL_FILE = False       # loader reads test_item_pool.csv
SHAPE = (6040, 3706, 1000000)          # (users, items, train interactions) of the dataset this config was tuned on


def main(device, model_config: dict, train_config: dict, evaluate_config: dict, data_loader, random_seed: int,
         silent: bool = False, auto: bool = False, query: bool = True):
    return _common.run_main(True, device, model_config, train_config, evaluate_config, data_loader,
                            random_seed, silent=silent, auto=auto, query=query, metric_list=METRIC_LIST,
                            use_item_pool=USE_ITEM_POOL)


if __name__ == '__main__':
    _common.cli(sys.modules[__name__], implicit=True, shape=SHAPE)
