"""reference global_config.py:1-2, overridable: INVPREF_DATASET_PATH / INVPREF_RESULT_SAVE_PATH."""
import os

RESULT_SAVE_PATH = os.environ.get("INVPREF_RESULT_SAVE_PATH", os.path.join(os.getcwd(), "result_save_path"))
DATASET_PATH = os.environ.get("INVPREF_DATASET_PATH", os.path.join(os.getcwd(), "dataset"))
