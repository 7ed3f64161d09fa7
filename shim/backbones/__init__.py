"""The reference's backbones/ package paths, re-exporting the native backbones (one module per reference file)."""
