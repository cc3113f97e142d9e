"""TEST-ONLY paramz.core stand-in module (import-time only)."""
