"""Device-backed counterparts of DistantSpeech/doa (see DESIGN.md for the reference file:line map)."""
