"""Drop-in for the reference package `metrics` (the training tracker)."""
