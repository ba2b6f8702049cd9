"""Evaluation, logging, timing, checkpointing (SURVEY.md section 5 auxiliary subsystems)."""
from .evaluate import get_loss_n_accuracy
from .logging import MetricLogger
from .timers import PhaseTimer
from .checkpoint import save_checkpoint, load_checkpoint

__all__ = ["get_loss_n_accuracy", "MetricLogger", "PhaseTimer", "save_checkpoint", "load_checkpoint"]
