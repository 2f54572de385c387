"""B200-native (sm_100a) BiLSTM caption encoder + mention-span heads: drop-in for the neural hot path of
cmcervantes/ImageCaptionLearn_py (`nn_utils/core.py` + `nn_utils/data.py:load_batch`)."""
__version__ = "0.1.0"
