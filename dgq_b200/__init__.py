"""dgq_b200 -- B200 (sm_100a) kernels and host glue behind DGQ's quantized UNet forward path."""
__all__ = ["ops"]
