"""Import shim (test infrastructure only) for `pytorch_lightning`. No arithmetic."""
