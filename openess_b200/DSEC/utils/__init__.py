"""Mirror of DSEC/utils/ (event slicing)."""
