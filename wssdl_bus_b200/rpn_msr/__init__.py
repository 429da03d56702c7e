"""Mirror of the reference's ``rpn_msr`` package (proposal / anchor-target / proposal-target
layers and anchor generation)."""
