"""Drop-in `src` package: the reference's import paths (src.audio_processing,
src.feature_extraction, src.models) backed by libdspfront.so.  Put the directory that contains
this package (dsp_audioreclabs_b200/dropin) ahead of the reference checkout on sys.path and
run.py / ablation_study.py / train_model.py / compare_feature_methods.py run unchanged."""
