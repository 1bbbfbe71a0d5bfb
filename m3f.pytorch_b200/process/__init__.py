"""Mirror of the reference's `process/` scripts that sit on the hot path (log-Mel extraction)."""
