"""The few `sr_tools` helpers the model-handler boundary needs (reference: Code/sr_tools)."""
