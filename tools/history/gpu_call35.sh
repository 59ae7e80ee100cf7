#!/usr/bin/env bash
timeout 900 python tools/bench_batch.py cfg3 2>&1 | tail -6
