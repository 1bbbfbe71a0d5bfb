"""Host-side mirror of the reference's `models/` package for the hot path (same class names, constructor
signatures, forward semantics and state_dict keys), executing on the sm_100a kernels of libm3t_b200.so."""
