"""groove_b200 — B200-native (sm_100a) implementation of the Transformer Groove Infilling training
and inference step, behind the reference's ``BaseGrooveTransformers`` module API."""
from .modules import GrooveTransformer, GrooveTransformerEncoder
from .evaluator import HVOMetrics, hvo_metrics_vector
from .training import FusedAdam, FusedSGD, calculate_loss, initialize_model, train_loop
from .pipeline import DeviceResidentLoader, HostBatchPrefetcher, HostPredictor
from .sweep import SweepPacker, params_from_config, sample_sweep

__all__ = ["GrooveTransformer", "GrooveTransformerEncoder", "calculate_loss", "initialize_model", "train_loop",
           "FusedSGD", "FusedAdam", "HVOMetrics", "hvo_metrics_vector", "SweepPacker", "sample_sweep", "params_from_config",
           "DeviceResidentLoader", "HostBatchPrefetcher", "HostPredictor"]
