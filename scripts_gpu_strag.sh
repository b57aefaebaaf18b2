#!/bin/bash
python tools/tail_probe2.py
