"""Drop-in for the reference's config.py: the same module constants and the same two
environment overrides (config.py:13,22), so experiments/run_experiments.py, train_model.py,
ablation_study.py and compare_feature_methods.py read identical values.

One deliberate difference: RESULTS_DIR is `DSP_RESULTS_DIR` if set, else `./results` under the
current working directory, instead of a directory next to this file (config.py:24 uses BASE_DIR);
like the reference (config.py:25-26) it is created at import time.  The name and meaning are unchanged.
"""
import os

BASE_DIR = os.path.dirname(os.path.abspath(__file__))

DATASET_TYPE = os.environ.get('DATASET_TYPE', 'name')
DATASET_PATHS = {
    'name': os.path.join(os.path.expanduser('~'), 'Downloads', 'speech_data_name'),
    'number': os.path.join(os.path.expanduser('~'), 'Downloads', 'speech_data_number'),
}
DATA_DIR = os.environ.get('SPEECH_DATA_DIR', DATASET_PATHS[DATASET_TYPE])

RESULTS_DIR = os.environ.get('DSP_RESULTS_DIR', os.path.join(os.getcwd(), 'results'))
os.makedirs(RESULTS_DIR, exist_ok=True)

SAMPLE_RATE = 44100
NORMALIZE = True

FRAME_LENGTH_MS = 25
FRAME_SHIFT_MS = 10
FRAME_LENGTH = int(SAMPLE_RATE * FRAME_LENGTH_MS / 1000)
FRAME_SHIFT = int(SAMPLE_RATE * FRAME_SHIFT_MS / 1000)

ENERGY_HIGH_RATIO = 0.5
ENERGY_LOW_RATIO = 0.1
ZCR_THRESHOLD_RATIO = 1.5

WINDOW_TYPES = ['rectangular', 'hamming', 'hanning']
FEATURE_STATS = ['mean', 'std', 'max', 'min', 'median']

KNN_N_NEIGHBORS = 3
SVM_C = 1.0
SVM_KERNEL = 'rbf'
MLP_HIDDEN_LAYERS = [64, 64, 32]
MLP_LEARNING_RATE = 0.005
MLP_EPOCHS = 1000
MLP_BATCH_SIZE = 108

TEST_SIZE = 0.2
RANDOM_SEED = 42

FIGURE_DPI = 150
FIGURE_SIZE = (12, 8)

LEARNING_RATES = [0.0001, 0.0003, 0.0005, 0.001, 0.003, 0.005, 0.008, 0.01, 0.03, 0.05, 0.08]
FRAME_LENGTH_MS_RANGE = [8, 10, 12, 15, 18, 20, 25, 30, 35, 40, 45, 50]
FRAME_SHIFT_MS_RANGE = [3, 5, 7, 8, 10, 12, 15, 18, 20, 25, 30]
