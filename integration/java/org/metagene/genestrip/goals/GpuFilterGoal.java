// GpuFilterGoal.java -- FilterGoal with the GPU filter behind it: makeFile (core/.../goals/FilterGoal.java:80-108) builds its
// FastqBloomFilter inline, so the subclass restates that method around GpuFastqBloomFilter.  The index is uploaded once per
// goal (first file) and released with the goal's last file.
package org.metagene.genestrip.goals;

import java.io.File;
import java.io.IOException;
import java.util.Map;

import org.metagene.genestrip.ExecutionContext;
import org.metagene.genestrip.GSConfigKey;
import org.metagene.genestrip.GSProject;
import org.metagene.genestrip.GSProject.GSFileType;
import org.metagene.genestrip.bloom.FastqBloomFilter;
import org.metagene.genestrip.bloom.GpuFastqBloomFilter;
import org.metagene.genestrip.gpu.GsNative;
import org.metagene.genestrip.io.StreamingResourceStream;
import org.metagene.genestrip.make.Goal;
import org.metagene.genestrip.make.ObjectGoal;

public class GpuFilterGoal<P extends GSProject> extends FilterGoal<P> {
    private final LoadIndexGoal<P> index;
    private final ExecutionContext bundle;
    private final long ctx;
    private long deviceFilter;

    @SafeVarargs
    public GpuFilterGoal(long ctx, P project, ObjectGoal<Map<String, StreamingResourceStream>, P> fastqMapGoal, LoadIndexGoal<P> indexedGoal,
            ExecutionContext bundle, Goal<P>... deps) {
        super(project, fastqMapGoal, indexedGoal, bundle, deps);
        this.ctx = ctx;
        this.index = indexedGoal;
        this.bundle = bundle;
    }

    @Override
    protected void makeFile(File file) throws IOException {
        FastqBloomFilter f = null;
        try {
            P project = getProject();
            StreamingResourceStream resources = fileToFastqs.get(file);
            File dumpFile = booleanConfigValue(GSConfigKey.WRITE_DUMPED_FASTQ)
                    ? project.getOutputFile("dumped", null, file.getName(), GSFileType.FASTQ_RES, isUseGZip())
                    : null;
            if (deviceFilter == 0) {
                deviceFilter = GpuFastqBloomFilter.upload(ctx, index.get());
            }
            f = new GpuFastqBloomFilter(deviceFilter, 1 << 20, 256L << 20, intConfigValue(GSConfigKey.KMER_SIZE), index.get(),
                    intConfigValue(GSConfigKey.MIN_POS_COUNT_FILTER), doubleConfigValue(GSConfigKey.POS_RATIO_FILTER),
                    intConfigValue(GSConfigKey.INITIAL_READ_SIZE_BYTES), intConfigValue(GSConfigKey.THREAD_QUEUE_SIZE), bundle,
                    booleanConfigValue(GSConfigKey.WITH_PROBS)) {
                @Override
                protected boolean isProgressBar() {
                    return booleanConfigValue(GSConfigKey.PROGRESS_BAR);
                }

                @Override
                protected String getProgressBarTaskName() {
                    return getKey().getName();
                }
            };
            f.runFilter(resources, file, dumpFile);
        } finally {
            if (f != null) {
                f.dump();
            }
        }
    }

    /** Releases the device copy of the index (call from the maker's dumpAll, core/.../GSMaker.java:75-80). */
    public void releaseDevice() {
        if (deviceFilter != 0) {
            GsNative.filterDestroy(deviceFilter);
            deviceFilter = 0;
        }
    }
}
