"""Device-backed counterparts of DistantSpeech/postfilter (see DESIGN.md for the reference file:line map)."""
