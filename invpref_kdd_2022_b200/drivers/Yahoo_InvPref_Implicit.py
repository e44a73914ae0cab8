"""Yahoo_InvPref_Implicit: InvPrefImplicit with the hyper-parameters of the reference driver
(reference Yahoo_InvPref_Implicit.py:17-67).  Run: ``python -m invpref_kdd_2022_b200.drivers.Yahoo_InvPref_Implicit [--epochs N] [--synthetic]``."""
import sys

from . import _common

MODEL_CONFIG: dict = {'env_num': 2, 'factor_num': 40, 'reg_only_embed': True, 'reg_env_embed': False}

TRAIN_CONFIG: dict = {'batch_size': 8192,
 'epochs': 1000,
 'cluster_interval': 5,
 'evaluate_interval': 10,
 'lr': 0.005,
 'invariant_coe': 3.351991776096847,
 'env_aware_coe': 9.988658447411407,
 'env_coe': 9.06447753571379,
 'L2_coe': 3.1351402017943117,
 'L1_coe': 0.4935216278026648,
 'alpha': 1.9053711444718746,
 'use_class_re_weight': True,
 'use_recommend_re_weight': False,
 'test_begin_epoch': 0,
 'begin_cluster_epoch': None,
 'stop_cluster_epoch': None}

EVALUATE_CONFIG: dict = {'top_k_list': [3, 5, 7], 'test_batch_size': 1024, 'eval_k': 5, 'eval_metric': 'ndcg'}

RANDOM_SEED_LIST = [17373331, 17373511, 17373423]

DATASET_PATH = '/Yahoo_all_data/'
METRIC_LIST = ['ndcg', 'recall', 'precision']
USE_ITEM_POOL = True            # Yahoo_InvPref_Implicit.py:87, :207: evaluator ranks only the test item pool
HAS_ITEM_POOL_FILE = True       # loader reads test_item_pool.csv
SHAPE = (15400, 1000, 250154)          # (users, items, train interactions) of the dataset this config was tuned on


def main(device, model_config: dict, train_config: dict, evaluate_config: dict, data_loader, random_seed: int,
         silent: bool = False, auto: bool = False, query: bool = True):
    return _common.run_main(True, device, model_config, train_config, evaluate_config, data_loader,
                            random_seed, silent=silent, auto=auto, query=query, metric_list=METRIC_LIST,
                            use_item_pool=USE_ITEM_POOL)


if __name__ == '__main__':
    _common.cli(sys.modules[__name__], implicit=True, shape=SHAPE)
