"""Drop-in ``basicsr`` namespace for the Shift-Net inference hot path (only the arch registry is provided;
training, datasets, losses and metrics of the reference's basicsr are out of scope -- SURVEY.md section 2)."""
