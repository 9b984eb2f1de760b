python tools/profile_train_step.py --shapes 2>/dev/null | tail -45
