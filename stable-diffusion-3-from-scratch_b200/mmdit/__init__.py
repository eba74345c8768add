"""Host-side binding and autograd glue for the B200-native MMDiT kernels."""
