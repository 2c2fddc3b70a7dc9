"""Drop-in replacement of the reference's `open_loop_training/ops` package."""
