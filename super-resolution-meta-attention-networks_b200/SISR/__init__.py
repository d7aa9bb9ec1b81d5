"""Host-side mirror of the reference's `SISR` package for the Deep-FIR hot path: same registry, handler
classes, method signatures and checkpoint layout (reference: Code/SISR), with the Q-model networks
dispatched to the B200 kernels (deepfir_b200)."""
