set -x
git stash -q 2>/dev/null
