"""Host-side mirror of the reference's `quant` package (the drop-in boundary)."""
