"""UNet graphs (SD v1.4 and SDXL) with the reference's module tree; see common.py."""
