"""Drop-in entry points mirroring activate_trained_model/test_imageTMO.py and test_videoTMO.py."""
