"""Device-backed counterparts of DistantSpeech/beamformer (see DESIGN.md for the reference file:line map)."""
