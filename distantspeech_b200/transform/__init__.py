"""Device-backed counterparts of DistantSpeech/transform (see DESIGN.md for the reference file:line map)."""
