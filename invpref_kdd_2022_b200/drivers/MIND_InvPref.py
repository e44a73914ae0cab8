"""MIND_InvPref: InvPrefImplicit with the hyper-parameters of the reference driver
(reference MIND_InvPref.py:17-67).  Run: ``python -m invpref_kdd_2022_b200.drivers.MIND_InvPref [--epochs N] [--synthetic]``."""
import sys

from . import _common

MODEL_CONFIG: dict = {'env_num': 6, 'factor_num': 40, 'reg_only_embed': True, 'reg_env_embed': False}

TRAIN_CONFIG: dict = {'batch_size': 262144,
 'epochs': 1000,
 'cluster_interval': 5,
 'evaluate_interval': 10,
 'lr': 0.001,
 'invariant_coe': 0.41343891722673093,
 'env_aware_coe': 9.833594297680568,
 'env_coe': 7.521558049068597,
 'L2_coe': 4.324061954456766,
 'L1_coe': 0.33322012936680223,
 'alpha': 1.5359474241627789,
 'use_class_re_weight': True,
 'use_recommend_re_weight': False,
 'test_begin_epoch': 0,
 'begin_cluster_epoch': None,
 'stop_cluster_epoch': None}

EVALUATE_CONFIG: dict = {'top_k_list': [5, 10, 20, 40], 'test_batch_size': 256, 'eval_k': 5, 'eval_metric': 'ndcg'}

RANDOM_SEED_LIST = [17373331, 17373511, 17373423]

DATASET_PATH = '/MIND_all_data/'
METRIC_LIST = ['ndcg', 'recall', 'precision']
USE_ITEM_POOL = True            # MIND_InvPref.py:87, :201: evaluator ranks only the test item pool
HAS_ITEM_POOL_FILE = True       # loader reads test_item_pool.csv
SHAPE = (50000, 51283, 4194304)          # (users, items, train interactions) of the dataset this config was tuned on


def main(device, model_config: dict, train_config: dict, evaluate_config: dict, data_loader, random_seed: int,
         silent: bool = False, auto: bool = False, query: bool = True):
    return _common.run_main(True, device, model_config, train_config, evaluate_config, data_loader,
                            random_seed, silent=silent, auto=auto, query=query, metric_list=METRIC_LIST,
                            use_item_pool=USE_ITEM_POOL)


if __name__ == '__main__':
    _common.cli(sys.modules[__name__], implicit=True, shape=SHAPE)
