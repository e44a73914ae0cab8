"""B200-native InvPref hot path: fused train step + EM environment re-assignment.

Host side in Python (mirrors the reference's ``models.py`` / ``train.py`` API), arithmetic in
``libinvpref_b200.so`` (hand-written CUDA for sm_100a behind the C ABI of ``include/invpref_b200.h``).
"""
__version__ = "0.1.0"
