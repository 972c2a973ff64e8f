#!/bin/bash
# default GPU job: the round's evidence capture (tests, bench lines, launch lists, full captures, sweep)
bash tools/capture_profiles.sh
