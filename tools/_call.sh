bash tools/collect_round.sh r02g "tests smoke bench"
