"""Host-side mirror of ``brevitas.core`` for the fake-quant hot path, on the B200 kernels."""
