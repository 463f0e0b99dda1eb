"""Namespace package for the B200-native BirdNest.Audio FLAC decode path (see birdnest.audio_b200)."""
