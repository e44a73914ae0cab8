"""Yahoo_InvPref_explicit: InvPrefExplicit with the hyper-parameters of the reference driver
(reference Yahoo_InvPref_explicit.py:17-67).  Run: ``python -m invpref_kdd_2022_b200.drivers.Yahoo_InvPref_explicit [--epochs N] [--synthetic]``."""
import sys

from . import _common

MODEL_CONFIG: dict = {'env_num': 5, 'factor_num': 40, 'reg_only_embed': True, 'reg_env_embed': False}

TRAIN_CONFIG: dict = {'batch_size': 131072,
 'epochs': 1000,
 'cluster_interval': 20,
 'evaluate_interval': 10,
 'lr': 0.001,
 'invariant_coe': 0.007375309563638757,
 'env_aware_coe': 7.207790368836971,
 'env_coe': 7.30272189219841,
 'L2_coe': 5.105587170019545,
 'L1_coe': 0.004098813161410509,
 'alpha': None,
 'use_class_re_weight': False,
 'use_recommend_re_weight': False,
 'test_begin_epoch': 0,
 'begin_cluster_epoch': None,
 'stop_cluster_epoch': None}

EVALUATE_CONFIG: dict = {'eval_metric': 'mse'}

RANDOM_SEED_LIST = [17373331, 17373511, 17373423]

DATASET_PATH = '/Yahoo_explicit_all_data/'
METRIC_LIST = ['mse', 'rmse', 'mae']
SHAPE = (15400, 1000, 311704)          # (users, items, train interactions) of the dataset this config was tuned on


def main(device, model_config: dict, train_config: dict, evaluate_config: dict, data_loader, random_seed: int,
         silent: bool = False, auto: bool = False, query: bool = True):
    return _common.run_main(False, device, model_config, train_config, evaluate_config, data_loader,
                            random_seed, silent=silent, auto=auto, query=query, metric_list=METRIC_LIST)


if __name__ == '__main__':
    _common.cli(sys.modules[__name__], implicit=False, shape=SHAPE)
