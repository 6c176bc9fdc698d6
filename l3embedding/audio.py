"""l3embedding/audio.py of the reference (pcm2float :4-31) -> l3embedding_b200.audio."""
from l3embedding_b200.audio import pcm2float  # noqa: F401
