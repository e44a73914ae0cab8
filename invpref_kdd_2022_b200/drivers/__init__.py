"""Runnable equivalents of the reference's five InvPref driver scripts (same module names, config dict
names/keys, ``main()`` signature and return value); device and dataset root are configurable instead of
hard-coded (reference ``global_config.py:1-2``, ``torch.cuda.set_device(N)`` at import)."""
