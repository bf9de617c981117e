"""Drop-in for the reference package `augmentation` (the SMPL / camera part only)."""
