#!/bin/bash
SKIP_NCU=1 bash tools/final_round.sh
