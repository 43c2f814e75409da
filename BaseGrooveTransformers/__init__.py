"""Drop-in alias: ``from BaseGrooveTransformers import initialize_model, calculate_loss, train_loop``
(the import the reference's train.py:12 and tutorial.py:6 use) resolves to the B200-native package."""
from transformergrooveinfilling_b200 import calculate_loss, initialize_model, train_loop

__all__ = ["initialize_model", "calculate_loss", "train_loop"]
