/* TEST INFRASTRUCTURE stub: nothing of kseq.h is used on this path */
