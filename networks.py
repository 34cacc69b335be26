"""Drop-in replacement for the reference's top-level `networks` module (networks.py:1-285).

Put this repository ahead of the reference on sys.path and regression/reg_task.py,
classification/class_task.py, reinforcement_learning/bandits.py, utils/load_model_utils.py and
weight_pruning.py resolve `from networks import ...` to the sm_100a implementation unchanged.
"""
import bnn_b200  # noqa: F401  (registers the package)
from bnn_b200 import (ScaleMixtureGaussian, GaussianNode, BayesianLinear, BayesianLinearLR,  # noqa: F401
                      BayesianNetwork, MLP, MLP_Dropout)
from config import *  # noqa: F401,F403  (the reference does the same, networks.py:12)
