"""Driver mirrors: one module per reference driver whose closures are in the device menu."""
