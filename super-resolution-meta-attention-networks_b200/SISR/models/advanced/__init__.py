"""Non-meta baselines (reference: Code/SISR/models/advanced/).  Only the handlers whose networks run on the B200 path
are mirrored: RCAN and EDSR."""
