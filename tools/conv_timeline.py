"""Wait-time accounting of conv_tc2's roles (experiments build: SYNT_EXPERIMENTS=1 python -m synt_isic_b200.build --force).
    SYNT_CONV_TL=1 python tools/conv_timeline.py      -> one "[conv_tl ...]" line per launch on stderr"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import exp_conv64 as e   # noqa: F401  (runs its table; the records are printed by the library)
