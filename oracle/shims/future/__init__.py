"""Test-infrastructure stub: stands in for the `future` package so that the read-only
reference tree under /root/reference can be imported in this container."""
