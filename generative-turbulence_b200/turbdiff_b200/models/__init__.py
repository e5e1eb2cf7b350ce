"""Host-side mirror of the reference's ``turbdiff.models`` interface for the denoising path."""
