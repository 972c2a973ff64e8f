#!/bin/bash
bash tools/capture_profiles.sh
