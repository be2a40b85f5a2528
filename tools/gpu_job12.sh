python -m pytest tests -m gpu -x -q -k "plan_steps" 2>&1 | tail -12
