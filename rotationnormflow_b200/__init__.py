"""rotationnormflow_b200 -- B200-native (sm_100a) hot path of RotationNormFlow: evaluating and inverting the
composed discrete normalizing flow on SO(3).  Drop-in for the reference's ``flow/flow.py`` API; the compute lives in
``librnf_b200.so`` (hand-written CUDA, C ABI in ``include/rnf_abi.h``).  No CPU fallback."""
from .config import load_config
from .flow import (ConditionalTransform, Condition16Trans, Condition16TransLU, ConditionLU, ConditionRot, Flow, MobiusFlow, Uncondition16Trans,
                   Uncondition16TransLU, UnconditionLU, UnconditionRot, get_affine, get_flow, get_mobius,
                   Condition9Trans, Uncondition9Trans, Uncondition9TransLU, Condition36Trans, Uncondition36Trans, Condition9RotL,
                   Uncondition9RotL, Condition9RotR, Uncondition9RotR, Condition9RotRSmith, Uncondition9RotRSmith)

__all__ = ["load_config", "get_flow", "Flow", "MobiusFlow", "ConditionalTransform", "Uncondition16Trans",
           "Uncondition16TransLU", "UnconditionLU", "Condition16Trans", "Condition16TransLU", "ConditionLU", "UnconditionRot", "ConditionRot",
           "get_affine", "get_mobius", "Condition9Trans", "Uncondition9Trans", "Uncondition9TransLU", "Condition36Trans",
           "Uncondition36Trans", "Condition9RotL", "Uncondition9RotL", "Condition9RotR", "Uncondition9RotR", "Condition9RotRSmith",
           "Uncondition9RotRSmith"]
