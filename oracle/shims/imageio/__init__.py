"""Stub of imageio (unused on the ConvBPDN path)."""
