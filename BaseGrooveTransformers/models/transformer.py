"""``from BaseGrooveTransformers.models.transformer import GrooveTransformerEncoder, GrooveTransformer``"""
from transformergrooveinfilling_b200.modules import GrooveTransformer, GrooveTransformerEncoder

__all__ = ["GrooveTransformerEncoder", "GrooveTransformer"]
