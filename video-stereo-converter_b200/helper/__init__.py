"""Drop-in package shadowing the reference's `helper` for the SBS stage: put the directory that
contains this package ahead of the reference tree on sys.path (or copy `helper/stereo_core.py` over the
reference's file) and `sbs_generator.py`, `sbs_tester.py` and `helper/config_manager.py` import the
B200 implementation through their unchanged `from helper.stereo_core import ...` lines."""
