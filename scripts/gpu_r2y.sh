#!/usr/bin/env bash
exec bash scripts/gpu_round_end.sh
