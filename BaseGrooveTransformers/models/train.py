from transformergrooveinfilling_b200.training import calculate_loss, initialize_model, train_loop

__all__ = ["initialize_model", "calculate_loss", "train_loop"]
