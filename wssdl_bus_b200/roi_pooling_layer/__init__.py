"""Mirror of the reference's ``roi_pooling_layer`` package."""
